// azb_nabla.hpp — C++ host-side mirror of the reference's optimizer surface for the c21 path, header-only, on top
// of the C ABI in azb.h.  The reference's host is Rust (nightly; no toolchain in this image), so the compiled host
// layer above the ABI is C++; INTEGRATION.md shows the Rust `extern "C"` binding a maintainer would add.
//
// Mirrors: NablaOptimizer (az-discrete-opt/src/nabla/optimizer/mod.rs:7-22,39-363), NablaModel
// (nabla/model/mod.rs:4-8), ArgminData (log.rs:1-11), ArgminImprovement (optimizer/mod.rs:24-27).
// Error behaviour: the reference panics (process abort); here every violation throws azb::Error.
#pragma once
#include <cstdint>
#include <functional>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "azb.h"

namespace azb {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

struct ArgminData {  // log.rs:1-11
    std::vector<uint8_t> parents;
    std::vector<uint32_t> permitted;
    double lambda_1 = 0;
    uint32_t mu = 0;
    float eval = 0;
};

// NablaModel (nabla/model/mod.rs:4-8) for models that live on the host
struct NablaModel {
    virtual ~NablaModel() = default;
    virtual void write_predictions(const std::vector<float> &states, std::vector<float> &predictions) = 0;
    // nabla/model/mod.rs:7: returns the loss
    virtual float update_model(const std::vector<float> & /*states*/, const std::vector<float> & /*observations*/,
                               const std::vector<float> & /*action_weights*/) { return 0.f; }
};
// nabla/model/mod.rs:10-23
struct TrivialModel : NablaModel {
    void write_predictions(const std::vector<float> &, std::vector<float> &) override {}
};

class NablaOptimizer {
   public:
    // par_new (optimizer/mod.rs:39-118).  model == nullptr selects the built-in device MLP (ActionModel,
    // nabla/model/dfdx.rs:18-53) initialised from mlp_seed.
    static NablaOptimizer par_new(const azb_config &cfg, const std::vector<uint8_t> &root_parents,
                                  const std::vector<uint32_t> &root_permitted, NablaModel *model, uint64_t mlp_seed = 0) {
        NablaOptimizer o;
        o.cfg_ = cfg;
        o.model_ = model;
        o.cfg_.prior_mode = model ? AZB_PRIOR_INJECTED : AZB_PRIOR_MLP;
        o.ck(azb_create(&o.cfg_, &o.h_));
        o.ck(azb_get_config(o.h_, &o.cfg_));
        const uint32_t n = o.cfg_.n_vertices;
        o.a_ = (n - 1) * (n - 2) / 2 - 1;
        o.w_ = (o.a_ + 31) / 32;
        o.state_vecs_.assign((size_t)o.cfg_.n_roots * 2 * o.a_, 0.f);
        o.h_theta_.assign((size_t)o.cfg_.n_roots * o.a_, 0.f);
        if (!model) o.ck(azb_mlp_init(o.h_, mlp_seed));
        o.reseed(root_parents, root_permitted);
        return o;
    }
    NablaOptimizer(NablaOptimizer &&o) noexcept { *this = std::move(o); }
    NablaOptimizer &operator=(NablaOptimizer &&o) noexcept {
        std::swap(h_, o.h_);
        cfg_ = o.cfg_;
        model_ = o.model_;
        a_ = o.a_;
        w_ = o.w_;
        state_vecs_.swap(o.state_vecs_);
        h_theta_.swap(o.h_theta_);
        return *this;
    }
    ~NablaOptimizer() {
        if (h_) azb_destroy(h_);
    }

    // par_roll_out_episodes (optimizer/mod.rs:121-191): one batched step; the value is ArgminImprovement::Improved
    std::optional<ArgminData> par_roll_out_episodes() {
        bool improved = false;
        if (!model_) {
            uint32_t n = 0;
            ck(azb_step(h_, 1, nullptr, 0, &n));
            improved = n > 0;
        } else {
            ck(azb_rollout_host(h_, state_vecs_.data()));
            model_->write_predictions(state_vecs_, h_theta_);
            int imp = 0;
            ck(azb_add_actions_host(h_, h_theta_.data(), &imp));
            improved = imp != 0;
        }
        if (!improved) return std::nullopt;
        return argmin_data();
    }
    // the example's inner loop `for episode in 1..=episodes` (04-c21-tree.rs:142-150), fused on the device
    std::vector<azb_improvement> roll_out(uint32_t n_steps) {
        std::vector<azb_improvement> log(n_steps ? n_steps : 1);
        uint32_t n = 0;
        ck(azb_step(h_, n_steps, log.data(), (uint32_t)log.size(), &n));
        log.resize(n < log.size() ? n : log.size());
        return log;
    }
    // The example's per-step loop (04-c21-tree.rs:140-150) with the trees running ahead of the caller: on_step(step,
    // improved, record) for each of the n_steps as soon as every tree has finished that step, then the batch is
    // completed, so argmin_data() afterwards is what n_steps calls of par_roll_out_episodes would have left.
    template <class F>
    void roll_out_ahead(uint32_t n_steps, F &&on_step) {
        ck(azb_step_enqueue(h_, n_steps));
        for (uint32_t s = 0; s < n_steps; ++s) {
            azb_improvement rec{};
            int improved = 0;
            ck(azb_step_poll(h_, &rec, &improved));
            on_step(rec.step, improved != 0, rec);
        }
        uint32_t n = 0;
        ck(azb_step(h_, 0, nullptr, 0, &n));
    }
    ArgminData argmin_data() {  // optimizer/mod.rs:361-363
        ArgminData d;
        d.parents.resize(cfg_.n_vertices);
        d.permitted.resize(w_);
        ck(azb_get_argmin(h_, d.parents.data(), d.permitted.data(), &d.lambda_1, &d.mu, &d.eval));
        return d;
    }
    // the inputs par_update_model hands to NablaModel::update_model (optimizer/mod.rs:253-280)
    void observations(uint32_t n_obs_tol, std::vector<float> &state_vecs, std::vector<float> &obs, std::vector<float> &w) {
        state_vecs.resize((size_t)cfg_.n_roots * 2 * a_);
        obs.resize((size_t)cfg_.n_roots * a_);
        w.resize(obs.size());
        ck(azb_write_observations(h_, n_obs_tol, state_vecs.data(), obs.data(), w.data()));
    }
    // par_update_model (optimizer/mod.rs:249-281) with the built-in ActionModel: root vectors, observations, f32
    // forward/backward and the Adam step on the device (all-reduced if a communicator is attached); returns the loss
    float par_update_model(uint32_t n_obs_tol) {
        float loss = 0.f;
        if (!model_) {
            ck(azb_update_model(h_, n_obs_tol, &loss));
        } else {
            std::vector<float> sv, obs, w;
            observations(n_obs_tol, sv, obs, w);
            loss = model_->update_model(sv, obs, w);
        }
        return loss;
    }
    // par_reset_trees (optimizer/mod.rs:284-360) with the example's modify_root policy (04-c21-tree.rs:172-206) on the
    // device; k_min/k_max = num_permitted_actions_range (0,0 = 5..=ACTION/2)
    void par_reset_trees(uint64_t seed, uint32_t k_min = 0, uint32_t k_max = 0) { ck(azb_reset_trees(h_, seed, k_min, k_max)); }
    // par_reset_trees (optimizer/mod.rs:284-360) with the caller's modify_root already applied to the roots
    void par_reset_trees(const std::vector<uint8_t> &root_parents, const std::vector<uint32_t> &root_permitted) {
        reseed(root_parents, root_permitted);
    }
    azb_handle *handle() { return h_; }
    const azb_config &config() const { return cfg_; }
    uint32_t action_dim() const { return a_; }

   private:
    NablaOptimizer() = default;
    void reseed(const std::vector<uint8_t> &p, const std::vector<uint32_t> &m) {
        if (p.size() != (size_t)cfg_.n_roots * cfg_.n_vertices || m.size() != (size_t)cfg_.n_roots * w_)
            throw Error(AZB_ERR_INVALID, "root array sizes");
        ck(azb_set_roots(h_, p.data(), m.data()));
        if (model_) {  // one model call on the root vectors (optimizer/mod.rs:65-72)
            const uint32_t n = cfg_.n_vertices;
            std::fill(state_vecs_.begin(), state_vecs_.end(), 0.f);
            for (uint32_t i = 0; i < cfg_.n_roots; ++i) {
                float *v = state_vecs_.data() + (size_t)i * 2 * a_;
                for (uint32_t c = 2; c + 1 < n; ++c) v[c * (c - 1) / 2 + p[(size_t)i * n + c] - 1] = 1.f;
                for (uint32_t a = 0; a < a_; ++a)
                    if ((m[(size_t)i * w_ + (a >> 5)] >> (a & 31)) & 1u) v[a_ + a] = 1.f;
            }
            std::fill(h_theta_.begin(), h_theta_.end(), 0.f);
            model_->write_predictions(state_vecs_, h_theta_);
            ck(azb_set_priors(h_, h_theta_.data()));
        }
        ck(azb_init_trees(h_));
    }
    void ck(int rc) {
        if (rc != AZB_OK)
            throw Error(rc, std::string(azb_strerror(rc)) + ": " + (h_ ? azb_last_error(h_) : ""));
    }
    azb_handle *h_ = nullptr;
    azb_config cfg_{};
    NablaModel *model_ = nullptr;
    uint32_t a_ = 0, w_ = 0;
    std::vector<float> state_vecs_, h_theta_;
};

// ---- ConnectedBitsetGraph<N, B32> (simple_graph/connected_bitset_graph/mod.rs:18-21) ----
template <uint32_t N>
struct ConnectedBitsetGraph {
    uint32_t neighborhoods[N];  // bit u of neighborhoods[v] <=> uv is an edge
};
struct Conjecture2Dot1Cost {  // mod.rs:340-344; only matching.len() is read anywhere, so the size is what is kept
    double lambda_1;
    uint32_t matching_number;
};
// conjecture_2_1_cost (:319-337) for a batch; `action_kinds` (may be null) receives action_kinds() (:134-154) per graph
// as N (N - 1) bits in AddOrDeleteEdge::action_index order (bitset_graph/space/action.rs:10-19), padded to u32 words
template <uint32_t N>
std::vector<Conjecture2Dot1Cost> conjecture_2_1_costs(NablaOptimizer &on, const std::vector<ConnectedBitsetGraph<N>> &graphs,
                                                      std::vector<uint32_t> *action_kinds = nullptr) {
    static_assert(N >= 2 && N <= 32, "B32 neighbourhoods");
    const uint32_t m = (uint32_t)graphs.size(), kw = (N * (N - 1) + 31) / 32;
    std::vector<double> l1(m);
    std::vector<uint32_t> mu(m);
    if (action_kinds) action_kinds->assign((size_t)m * kw, 0u);
    const int rc = m ? azb_eval_graph_costs(on.handle(), &graphs[0].neighborhoods[0], m, N, l1.data(), mu.data(),
                                            action_kinds ? action_kinds->data() : nullptr, nullptr)
                     : AZB_OK;
    if (rc != AZB_OK) throw Error(rc, std::string(azb_strerror(rc)) + ": " + azb_last_error(on.handle()));
    std::vector<Conjecture2Dot1Cost> out(m);
    for (uint32_t i = 0; i < m; ++i) out[i] = {l1[i], mu[i]};
    return out;
}

}  // namespace azb
