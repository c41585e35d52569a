/* ORACLE — test infrastructure, not product code.
 * Plain-C restatement of pred_to_count (workoutdetector/utils/inference_count.py:114-165), batched over videos with
 * the same array signature as wd_count_reps (include/wd_b200.h) so tests can compare buffers byte for byte.
 * Built by oracle/Makefile into oracle/_build/libcount_oracle.so. */
#include <stdint.h>

int oracle_count_reps(const int32_t* states, const int32_t* lens, int V, int Wmax, int step, int32_t* counts,
                      int32_t* reps, int reps_stride, int32_t* reps_len) {
    for (int v = 0; v < V; ++v) {
        const int32_t* p = states + (int64_t)v * Wmax;
        int len = lens ? lens[v] : Wmax;
        if (len > Wmax) len = Wmax;
        int count = 0, have_last = 0, last = 0, start_idx = 0;
        for (int idx = 0; idx < len; ++idx) {
            const int pred = p[idx];
            if (pred == -1) continue;                                     /* :151-152 */
            if (have_last && last != pred) {                              /* :154 */
                /* Python's pred % 2 == 1 is true for every odd integer, negative ones included */
                if ((pred & 1) == 1 && last == pred - 1) {                /* :155 */
                    if (reps && 2 * count + 1 < reps_stride) {
                        reps[(int64_t)v * reps_stride + 2 * count] = start_idx * step;     /* :157 */
                        reps[(int64_t)v * reps_stride + 2 * count + 1] = idx * step;       /* :158 */
                    }
                    ++count;
                }
            }
            last = pred;                                                  /* :159 */
            have_last = 1;
            if (pred != p[start_idx]) start_idx = idx;                    /* :160-162 */
        }
        counts[v] = count;
        if (reps_len) reps_len[v] = 2 * count;
    }
    return 0;
}

/* The 7-frame majority vote of count_by_image_model (utils/inference_count.py:213-224), same array signature as
 * wd_vote_states: states[v][f] = (sum of labels[v][max(0, f-window+1) .. f] >= votes), -1 past lens[v]. */
int oracle_vote_states(const int32_t* labels, const int32_t* lens, int V, int Fmax, int window, int votes,
                       int32_t* states) {
    for (int v = 0; v < V; ++v) {
        int len = lens ? lens[v] : Fmax;
        if (len > Fmax) len = Fmax;
        for (int f = 0; f < Fmax; ++f) {
            if (f >= len) {
                states[(int64_t)v * Fmax + f] = -1;
                continue;
            }
            int sum = 0;
            for (int j = f - window + 1 < 0 ? 0 : f - window + 1; j <= f; ++j) sum += labels[(int64_t)v * Fmax + j];
            states[(int64_t)v * Fmax + f] = sum >= votes ? 1 : 0;
        }
    }
    return 0;
}
