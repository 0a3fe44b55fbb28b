#!/usr/bin/env python
"""The reference's example loop (graph-state/examples/04-c21-tree.rs:76-208) on the B200 path, through the Python
mirror of NablaOptimizer (azdopt_b200/nabla.py) — same constants, same outputs:

    python examples/c21_epoch.py [--batch 512] [--epochs 3] [--episodes 800] [--out out_dir]

Per epoch: `episodes` search steps fused on the device (improvements are printed with their step like the example's
process_argmin), par_update_model (loss), par_reset_trees with the example's modify_root policy.  Writes a TensorBoard
event file with the example's scalar tags, the best tree as graph6 and the first search DAG as Graphviz DOT.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from azdopt_b200 import capi, nabla, observe  # noqa: E402

N = 19                       # 04-c21-tree.rs:33
C_LOWER, C_UPPER = 2.0, 15.0  # :58-68 for N = 19


def squish(c):  # :70-74
    return (c - C_LOWER) / (C_UPPER - C_LOWER)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)       # :54
    ap.add_argument("--epochs", type=int, default=3)        # :133 (250 in the example)
    ap.add_argument("--episodes", type=int, default=800)    # :134
    ap.add_argument("--n-obs-tol", type=int, default=200)   # :135
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default="c21_out")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    space = nabla.ROTModifyParentsOnce(n=N)
    model = nabla.ActionModel(seed=args.seed + 1, arithmetic="tc")
    roots = capi.generate_roots(args.seed, 0, args.batch, N)  # ROTWithActionPermissions::generate, k ~ U{5..A/2} (:85,108-112)
    opt = nabla.NablaOptimizer.par_new(space, roots, model, args.batch, max_steps=args.episodes)
    goal = squish(5.2)                                         # :117
    with open(os.path.join(args.out, "tfevents-losses"), "wb") as f:
        writer = observe.TensorboardWriter(f)
        writer.write_file_version()

        def process_argmin(a, step):                           # :119-130
            print(f"{a.eval:12.6f}\tlambda_1={a.lambda_1:.6f} mu={a.mu}")
            writer.write_cost(step, a.lambda_1, a.mu)
            writer.flush()
            return a.eval < goal

        process_argmin(opt.argmin_data(), 0)
        for epoch in range(1, args.epochs + 1):
            print(f"==== EPOCH: {epoch} ====")
            improved = opt.roll_out(args.episodes)             # the inner loop :141-150, fused on the device
            for (step, tree, node, ev) in improved:
                print(f"  step {args.episodes * (epoch - 1) + step + 1}: tree {tree} node {node} eval {ev:.6f}")
            if improved and process_argmin(opt.argmin_data(), args.episodes * epoch):
                print("state is optimal:", opt.argmin_data().parents.tolist())
                break
            loss = opt.par_update_model(args.n_obs_tol)        # :163
            writer.write_loss(args.episodes * epoch, loss)
            writer.write_cost(args.episodes * epoch, opt.argmin_data().lambda_1, opt.argmin_data().mu)
            writer.flush()
            print(f"  loss {loss:.6f}")
            with open(os.path.join(args.out, "tree.dot"), "w") as g:   # :153-157 (the example renders PNGs)
                g.write(opt.tree_dot(0))
            opt.par_reset_trees(seed=args.seed + 7)            # :172-207
    with open(os.path.join(args.out, "best.g6"), "wb") as g:
        g.write(opt.argmin_graph6() + b"\n")
    print("best tree (graph6):", opt.argmin_graph6().decode())
    opt.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
