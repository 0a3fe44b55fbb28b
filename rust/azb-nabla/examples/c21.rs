//! graph-state/examples/04-c21-tree.rs with the optimizer swapped: the same constants, closures and loop shape
//! (`:33-54,58-74,96-115,131-208`), `NablaOptimizer` -> `B200Optimizer`, the dfdx `ActionModel` -> the built-in MLP
//! (same architecture and Adam settings, `:42-52,87-92`).  Tensorboard output is left out.
#![feature(isqrt)]

use az_discrete_opt::{log::ArgminData, nabla::{model::TrivialModel, optimizer::ArgminImprovement}};
use azb_nabla::{B200Optimizer, Cost, Model, Space, State};
use graph_state::{rooted_tree::modify_parent_once::ROTWithActionPermissions, simple_graph::connected_bitset_graph::Conjecture2Dot1Cost};
use rand::Rng;

const N: usize = 19;
const STATE: usize = (N - 1) * (N - 2) - 2;
const ACTION: usize = (N - 1) * (N - 2) / 2 - 1;
const HIDDEN: [usize; 3] = [512, 1024, 512];
const BATCH: usize = 4096; // the reference ships 512 (:54); BASELINE configs[1] is 4096 roots on one B200

const C_LOWER_BOUND: usize = 2;
const C_UPPER_BOUND: usize = {
    const _SQRT: usize = (N - 1).isqrt();
    const SQRT: usize = if _SQRT * _SQRT == N - 1 { _SQRT } else { _SQRT + 1 };
    const MU_MAX: usize = (N + 1) / 2;
    SQRT + MU_MAX
};

fn squish(x: f32) -> f32 {
    const SLOPE: f32 = 1.0 / ((C_UPPER_BOUND - C_LOWER_BOUND) as f32);
    let x = x - C_LOWER_BOUND as f32;
    SLOPE * x
}

/// dfdx's default `Linear` initialisation, in dfdx parameter order: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
fn init_params(rng: &mut impl Rng) -> Vec<f32> {
    let dims = [STATE, HIDDEN[0], HIDDEN[1], HIDDEN[2], ACTION];
    let mut p = Vec::new();
    for l in 0..4 {
        let bound = 1.0 / (dims[l] as f32).sqrt();
        for _ in 0..dims[l] * dims[l + 1] + dims[l + 1] {
            p.push(rng.gen_range(-bound..bound));
        }
    }
    p
}

fn main() -> eyre::Result<()> {
    let num_permitted_actions_range = 5..=(ACTION / 2);
    const SPACE: Space<N> = Space::new(
        |t| t.conjecture_2_1_cost(),
        |c| {
            let Conjecture2Dot1Cost { matching, lambda_1 } = c;
            squish(matching.len() as f32 + *lambda_1 as f32)
        },
        |c_s, h_theta_sa| c_s - h_theta_sa,
        |_c_s, c_as_star| c_as_star,
    );
    let episodes: u32 = 800;
    let epochs: usize = 250;
    let n_obs_tol = 200;
    let n_as_tol = |len: usize| -> u32 { [200, 50, 50].get(len).copied().unwrap_or(25) };
    let model = Model::Device { params: init_params(&mut rand::thread_rng()), tensor_cores: true };
    let mut optimizer: B200Optimizer<N, TrivialModel> = B200Optimizer::par_new(
        SPACE,
        || {
            let mut rng = rand::thread_rng();
            let num_permitted_actions = rng.gen_range(num_permitted_actions_range.clone());
            ROTWithActionPermissions::generate(&mut rng, num_permitted_actions)
        },
        model,
        BATCH,
        n_as_tol,
        episodes,
    );
    let goal: f32 = squish(5.2);
    let process_argmin = |argmin: &ArgminData<State<N>, Cost>| -> eyre::Result<()> {
        let ArgminData { state, cost, eval } = argmin;
        println!("{eval:12}\t{cost:?}");
        if *eval < goal { Err(eyre::eyre!(format!("state is optimal:\n{state}"))) } else { Ok(()) }
    };
    process_argmin(optimizer.argmin_data())?;
    for epoch in 1..=epochs {
        println!("==== EPOCH: {epoch} ====");
        optimizer.roll_out_ahead(episodes); // the trees run ahead; the loop below keeps the reference's shape
        for _episode in 1..=episodes {
            match optimizer.par_roll_out_episodes(n_as_tol) {
                ArgminImprovement::Improved(argmin) => process_argmin(argmin)?,
                ArgminImprovement::Unchanged => {}
            };
        }
        let loss = optimizer.par_update_model(n_obs_tol);
        println!("loss: {loss}");
        optimizer.par_reset_trees_c21(*num_permitted_actions_range.start() as u32, *num_permitted_actions_range.end() as u32);
    }
    Ok(())
}
