/* TEST INFRASTRUCTURE -- plain C restatement of the reference's numba kernels (the loops that are
 * loops in the reference).  Built by `make -C oracle` into oracle/liboracle_c.so; loaded only by
 * tests/ and by bench.py's CPU-baseline leg.  Never linked into the product library.
 *
 * Each function cites the reference code it follows (paths under /root/reference).
 * Compile WITHOUT -ffast-math / FMA contraction so the f64 results match numba (-ffp-contract=off).
 */
#include <stdint.h>
#include <string.h>

/* numba `_gae` -- tianshou/algorithm/algorithm_base.py:1085-1140 (v_s / v_s_ given as f64). */
void oracle_gae(const double* v_s, const double* v_s_, const double* rew, const uint8_t* end_flag,
                int64_t n, double gamma, double lam, double* out) {
    double gae = 0.0;
    const double gl = gamma * lam;
    for (int64_t i = n - 1; i >= 0; --i) {
        const double delta = rew[i] + v_s_[i] * gamma - v_s[i];
        const double discount = (1.0 - (double)(end_flag[i] != 0)) * gl;
        gae = delta + discount * gae;
        out[i] = gae;
    }
}

/* numba `_nstep_return` -- algorithm_base.py:1160-1222. idx is [n_step][I], target_q [I][A] f32. */
void oracle_nstep_return(const double* rew, const uint8_t* end_flag, const float* target_q,
                         const int64_t* idx, int64_t I, int64_t A, int32_t n_step, double gamma,
                         double* out) {
    for (int64_t i = 0; i < I; ++i) {
        double acc = 0.0;
        int gammas = n_step;
        for (int n = n_step - 1; n >= 0; --n) {
            const int64_t now = idx[(int64_t)n * I + i];
            if (end_flag[now]) { gammas = n + 1; acc = 0.0; }
            acc = rew[now] + gamma * acc;
        }
        double gp = 1.0;
        for (int k = 0; k < gammas; ++k) gp = gp * gamma;
        for (int64_t a = 0; a < A; ++a) out[i * A + a] = (double)target_q[i * A + a] * gp + acc;
    }
}

static int64_t pymod(int64_t a, int64_t m) { int64_t r = a % m; return r < 0 ? r + m : r; }

/* numba `_next_index` / `_prev_index` -- tianshou/data/buffer/manager.py:339-363 / :311-336. */
void oracle_next_index(const int64_t* index, int64_t n, const int64_t* offset, int64_t E,
                       const uint8_t* done, const int64_t* last_index, const int64_t* lengths,
                       int64_t* out) {
    for (int64_t t = 0; t < n; ++t) {
        const int64_t i = pymod(index[t], offset[E]);
        int64_t e = 0;
        while (!(offset[e] <= i && i < offset[e + 1])) ++e;
        const int64_t L = lengths[e] > 1 ? lengths[e] : 1;
        const int64_t end = (done[i] != 0) | (i == last_index[e]);
        out[t] = pymod(i - offset[e] + 1 - end, L) + offset[e];
    }
}
void oracle_prev_index(const int64_t* index, int64_t n, const int64_t* offset, int64_t E,
                       const uint8_t* done, const int64_t* last_index, const int64_t* lengths,
                       int64_t* out) {
    for (int64_t t = 0; t < n; ++t) {
        const int64_t i = pymod(index[t], offset[E]);
        int64_t e = 0;
        while (!(offset[e] <= i && i < offset[e + 1])) ++e;
        const int64_t L = lengths[e] > 1 ? lengths[e] : 1;
        const int64_t sub = pymod(i - offset[e] - 1, L);
        const int64_t end = (done[sub + offset[e]] != 0) | (sub + offset[e] == last_index[e]);
        out[t] = pymod(sub + end, L) + offset[e];
    }
}

/* `_get_prefix_sum_idx` -- tianshou/data/utils/segtree.py:119-134. */
void oracle_prefix_sum_idx(const double* tree, int64_t bound, const double* value, int64_t n,
                           int64_t* out) {
    for (int64_t t = 0; t < n; ++t) {
        double v = value[t];
        int64_t index = 1;
        while (index < bound) {
            index *= 2;
            const double l = tree[index];
            if (l < v) { v -= l; index += 1; }
        }
        out[t] = index - bound;
    }
}
