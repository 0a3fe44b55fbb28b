// c21_epoch.cpp — the epoch loop of graph-state/examples/04-c21-tree.rs:106-208 over the C++ mirror
// (include/azb_nabla.hpp): par_new, `episodes` fused steps per epoch, argmin logging, observation pass.
// The whole loop runs on the device: fused search steps, the training step (par_update_model) and the example's
// modify_root policy (par_reset_trees).   build: g++ -O2 -std=c++17 -Iinclude examples/c21_epoch.cpp -Lazdopt_b200/lib -lazb -o c21_epoch
#include <cstdio>
#include <cstdlib>

#include "azb_nabla.hpp"

int main(int argc, char **argv) {
    const uint32_t n = 19, batch = argc > 1 ? (uint32_t)atoi(argv[1]) : 512;  // 04-c21-tree.rs:33,54
    const uint32_t epochs = argc > 2 ? (uint32_t)atoi(argv[2]) : 2, episodes = 800, n_obs_tol = 200;  // :133-135
    const bool per_step = argc > 3 && atoi(argv[3]) != 0;  // the example's loop shape: one result per step, as it completes
    azb_config cfg;
    if (azb_config_default(&cfg, n, batch) != AZB_OK) return 1;
    cfg.mlp_mode = AZB_MLP_TC;             // the built-in model on the tensor cores ...
    cfg.async_workers = AZB_ASYNC_AUTO;    // ... and the asynchronous search kernel wherever the batch is large enough
    const uint32_t a = (n - 1) * (n - 2) / 2 - 1, w = (a + 31) / 32;
    std::vector<uint8_t> parents((size_t)batch * n);
    std::vector<uint32_t> permitted((size_t)batch * w);
    try {
        azb_generate_roots(0, 0, batch, n, 5, a / 2, parents.data(), permitted.data());
        auto opt = azb::NablaOptimizer::par_new(cfg, parents, permitted, nullptr, /*mlp_seed=*/1);
        auto best = opt.argmin_data();
        printf("%12g\tlambda_1=%.6f mu=%u\n", best.eval, best.lambda_1, best.mu);
        const float goal = (5.2f - 2.0f) / 13.0f;  // squish(5.2), 04-c21-tree.rs:117
        for (uint32_t epoch = 1; epoch <= epochs; ++epoch) {
            if (per_step)
                opt.roll_out_ahead(episodes, [&](uint32_t step, bool improved, const azb_improvement &imp) {
                    if (improved) printf("epoch %u step %u: tree %u node %u eval %g\n", epoch, step, imp.tree, imp.node, imp.eval);
                });
            else
                for (const auto &imp : opt.roll_out(episodes))
                    printf("epoch %u step %u: tree %u node %u eval %g\n", epoch, imp.step, imp.tree, imp.node, imp.eval);
            best = opt.argmin_data();
            printf("%12g\tlambda_1=%.6f mu=%u\n", best.eval, best.lambda_1, best.mu);
            if (best.eval < goal) break;
            const float loss = opt.par_update_model(n_obs_tol);                  // 04-c21-tree.rs:163
            printf("epoch %u: loss %g\n", epoch, loss);
            opt.par_reset_trees(/*seed=*/1234);                                  // 04-c21-tree.rs:172-207
        }
    } catch (const azb::Error &e) {
        fprintf(stderr, "azb error %d: %s\n", e.code, e.what());
        return 2;
    }
    return 0;
}
