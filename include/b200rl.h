/* libb200rl -- C-ABI of the B200-native PPO2 / DQN learner hot path.
 *
 * The reference (openai/baselines) has NO FFI: its hot path is TF1 graph ops + numpy loops called from
 * Python.  Each entry point below therefore names the reference interface (file:line) whose arithmetic it
 * replaces; INTEGRATION.md shows the ctypes binding a baselines maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (torch tensors in this repo); no hidden
 *    allocation, no host synchronisation inside; `stream` is a cudaStream_t passed as void*.
 *  - return 0 on success, negative on error (see B200RL_ERR_*); b200rl_last_error() gives the text.
 *  - fp16 tensors are IEEE half; "ld*" are row pitches in ELEMENTS.
 *  - rollout arrays are time-major [T, N] (env contiguous); `src_idx` arrays hold buffer offsets t*N+e.
 */
#ifndef B200RL_H
#define B200RL_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define B200RL_OK 0
#define B200RL_ERR_ARG (-1)
#define B200RL_ERR_CUDA (-2)
#define B200RL_ERR_UNSUPPORTED (-3)
#define B200RL_ERR_DRIVER (-4)

/* GEMM epilogues */
#define B200RL_MODE_F16_ACT 0    /* C16 = act(alpha*acc + bias)                      (forward)            */
#define B200RL_MODE_F32_STORE 1  /* C32 = alpha*acc + bias                           (heads)              */
#define B200RL_MODE_F32_ATOMIC 2 /* C32 += alpha*acc (red.add.f32; split-K capable)  (weight gradients)   */
#define B200RL_MODE_F16_DACT 3   /* C16 = alpha*acc * act'(saved)                    (data gradients)     */
#define B200RL_MODE_F16_SHUFFLE 4 /* conv data gradient, pixel-shuffle scatter: row (n,i,j), col (py,px,c)
                                     -> dx[n, s*i+py, s*j+px, c] * act'(saved there)  (b200rl_conv_gemm only) */
#define B200RL_ACT_NONE 0
#define B200RL_ACT_RELU 1
#define B200RL_ACT_TANH 2

const char* b200rl_last_error(void);
int b200rl_version(void);

/* GAE(lambda) backward scan: baselines/ppo2/runner.py:53-65 (bit-exact, float64 carry).
 * dones[t] = done BEFORE step t (runner.py:34); last_dones = runner.dones after the last step.
 * variant: -1 auto, 0 register-prefetch kernel, 1 TMA-bulk pipelined kernel (needs N % 32 == 0). */
int b200rl_gae_scan(const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
                    const uint8_t* last_dones, float* advs, float* returns, int T, int N, double gamma, double lam,
                    int variant, void* stream);

/* fp16 x fp16 -> fp32 tcgen05 GEMM: tf.matmul a2c/utils.py:63; after im2col also tf.nn.conv2d a2c/utils.py:56
 * and their gradients (ppo2/model.py:102).
 *   mn_major = 0 : A[M,K] (lda), B[N,K] (ldb), C = A * B^T
 *   mn_major = 1 : A[K,M] (lda), B[K,N] (ldb), C = A^T * B   (reduction over rows; use split_k > 1)
 * max_ctas <= 0 : one persistent CTA per SM. */
int b200rl_gemm_f16(const void* A, const void* B, void* C, const float* bias, const void* saved, int M, int N, int K,
                    long long lda, long long ldb, long long ldc, long long ld_saved, int mn_major, int mode, int act,
                    float alpha, int split_k, int max_ctas, int rm_C, int rm_OW, int rm_Wg, const void* saved_bits,
                    void* stream);
/* saved_bits (MODE_F16_DACT with ReLU, optional): uint16 words, bit k of word (row*ld_saved + col)/16 set iff the saved
 * activation element (row, col + k) > 0 -- read instead of `saved` (16x less mask traffic; conv_shift_fwd emits it). */
/* rm_C > 0 (fp16 outputs only): output column pix*rm_C + c is stored at ((pix/rm_OW)*rm_Wg + pix%rm_OW)*rm_C + c,
 * i.e. a [.., OH, OW, C] row is scattered into a zero-bordered [.., Hg, Wg, C] grid (fc1 dgrad -> conv3's dY). */

/* Shift-GEMM convolutions (tf.nn.conv2d a2c/utils.py:56 + gradients): stride-1 conv over X[(n,y,x) rows, C]
 * (C = 64 or 128; strided convs arrive space-to-depth transformed).  Each X row is loaded into smem once; filter
 * tap t is the same buffer read through a descriptor shifted by shifts[t] rows.
 *   fwd  : out[map(n,y,x), :N] = act(sum_t X[m+shift_t] * W[:, t*C:(t+1)*C]^T + bias)   for y < vy, x < vx
 *          dact = 1: out = (sum_t ...) * act'(saved[smap(n,y,x)])  (data gradient: X = zero-bordered dY, shifts <= 0)
 *   omap / smap: {mode, sN, sY, sX, Cq, s}: 0 = n*sN+y*sY+x*sX+col; 1 = depth->space; 2 = space->depth
 *   wgrad: G[t*C + c, n] += alpha * sum_m X[m+shift_t, c] * dY[m, n]  (dY on X's grid, zero at invalid positions) */
int b200rl_conv_shift_fwd(const void* X, long long B, int Hg, int Wg, int C, const void* W, long long ldw, int N,
                          int taps, const int* shifts, int vy, int vx, void* out, const long long* omap,
                          const void* saved, const long long* smap, const float* bias, int act, int dact, float alpha,
                          const void* u8_x, const long long* u8_idx, int u8_H, int u8_W, int u8_C, int u8_s,
                          void* act_bits_out, const void* saved_bits, int kx, void* stream);
/* act_bits_out (forward, optional): uint16[numel(out)/16]; bit k of word e/16 is set iff out element e+k > 0 (e = the
 * element offset the output map produces, always a multiple of 16).  saved_bits (dact, optional): the same array for
 * the saved activation; it is read instead of `saved` (1 bit instead of 16 per element of backward HBM traffic).
 * u8_x != NULL (first layer): X is ignored; producer warps gather uint8 images u8_x[u8_idx[n], H, W, C], cast them
 * to fp16 and build the space-to-depth (factor u8_s) tile directly in shared memory (models.py:19, ppo2.py:165). */
int b200rl_conv_shift_wgrad(const void* X, long long rows, int C, const void* dY, int N, int taps, const int* shifts,
                            float* G, long long ldg, float alpha, float* gbias, float alpha_b, int max_ctas,
                            const void* u8_x, const long long* u8_idx, int u8_H, int u8_W, int u8_C, int u8_s,
                            int kx, void* stream);
/* kx > 1 ("x-fold", forward and wgrad): the filter is ky rows of kx horizontally adjacent taps (row shift of tap
 * (a, b) = a*Wg + b).  `shifts` then lists only the ky row shifts a*Wg, and the kx taps of a filter row ride in the
 * MMA's N dimension, so every X slab is fetched from shared memory once per filter row instead of once per tap.
 *   fwd  : W is [kx*N, ky*C] with row b*N + n, column a*C + c = filter tap (a, b), input channel c, output channel n
 *   wgrad: G keeps its [(a*kx + b)*C + c, n] row order. */   /* gbias != NULL: gbias[n] += alpha_b * sum_m dY[m, n] (fused) */

/* Implicit-GEMM convolution (tf.nn.conv2d a2c/utils.py:56 and its gradients): the A operand is read
 * straight from the NHWC fp16 activation x[B,H,W,C] by TMA im2col mode (C = 16, 32 or 64 channels per tap).
 *   kind 0: out[B*OH*OW, N] = patches(x) * Wt^T, Wt = [N, R*S*C] fp16 (ldb); modes F16_ACT / F16_DACT /
 *           F16_SHUFFLE (data gradient of a stride-s conv written through the sh_* geometry)
 *   kind 1: out[R*S*C, N] (fp32, ldc) += alpha * patches(x)^T * dz, dz = [B*OH*OW, N] fp16 (ldb); split_k >= 1 */
int b200rl_conv_gemm(const void* x, long long B, int H, int W, int C, int R, int S, int stride_h, int stride_w,
                     int pad_h, int pad_w, int OH, int OW, const void* Wt_or_dz, long long ldb, void* out,
                     long long ldc, const float* bias, const void* saved, long long ld_saved, int N, int kind,
                     int mode, int act, float alpha, int split_k, int sh_H, int sh_W, int sh_C, int sh_s,
                     void* stream);
/* fp16 weight operand for the pixel-shuffle data gradient: out[s*s*Cin, ceil(R/s)^2*Cout] from HWIO fp32 w */
int b200rl_dgrad_weights(const float* w, void* out, int R, int S, int Cin, int Cout, int s, long long ld,
                         void* stream);

/* conv lowering (tf.nn.conv2d NHWC, a2c/utils.py:37-56; SAME padding for tf.contrib convolution2d,
 * common/models.py:241).  src_is_u8 fuses tf.cast(uint8->float) of models.py:19 and, through src_idx,
 * the minibatch gather arr[mbinds] of ppo2/ppo2.py:165. */
int b200rl_im2col(const void* x, int src_is_u8, const long long* src_idx, void* cols, long long B, int H, int W,
                  int C, int rf, int stride, int same_pad, void* stream);
/* gather + uint8->fp16 + space-to-depth: out[b,Y,X,(dy*s+dx)*C+c] = x[src_idx[b], s*Y+dy, s*X+dx, c] */
/* Device half of VecFrameStack.step_wait (common/vec_env/vec_frame_stack.py:17-25): for N envs of `pixels` pixels,
 * out[n, p, :] = roll(prev[n, p, :], -1) (one channel, as the reference does); zeroed where news[n]; the last c
 * bytes = frame[n, p, :].
 * uint8 tensors: prev/out [N, pixels, nstack*c], frame [N, pixels, c], news [N].  out must not alias prev. */
int b200rl_frame_stack(const void* prev, const void* frame, const void* news, void* out, long long N, long long pixels,
                       int nstack, int c, void* stream);

int b200rl_s2d_gather(const void* x, const long long* src_idx, void* out, long long B, int H, int W, int C, int s,
                      void* stream);
int b200rl_col2im(const void* dcols, const void* saved, void* dx, long long B, int H, int W, int C, int rf,
                  int stride, int same_pad, int act, void* stream);
int b200rl_colsum(const void* dz, float* db, long long rows, int C, long long ld, float alpha, void* stream);

/* act path: PolicyWithValue.step common/policies.py:77-96; CategoricalPd.sample/neglogp
 * common/distributions.py:164-201; DiagGaussianPd :238-248.  noise == NULL -> counter-based Philox. */
int b200rl_cat_step(const float* logits, long long ld, int nA, const float* vpred, long long ldv,
                    const float* uniforms, unsigned long long seed, unsigned long long offset,
                    const unsigned long long* offset_dev, long long* actions, float* values, float* neglogp,
                    long long B, void* stream);
int b200rl_gauss_step(const float* mean, long long ld, const float* logstd, int d, const float* vpred,
                      long long ldv, const float* normals, unsigned long long seed, unsigned long long offset,
                      const unsigned long long* offset_dev, float* actions, float* values, float* neglogp,
                      long long B, void* stream);
/* Scalars that change between replays of a CUDA-graph-captured launch sequence live in device memory:
 *   *_dev arguments (offset_dev of the samplers, cliprange_dev of the losses, lr_t_dev of clip_adam), when non-NULL,
 *   override the by-value argument; set_scalars writes up to 4 floats from its own kernel arguments (no host staging
 *   buffer to race with); counter_add advances the sampler's stream position after each acting pass. */
int b200rl_set_scalars(float* dst, int n, float a, float b, float c, float d, void* stream);
/* Device minibatch shuffle: ppo2/ppo2.py:160 (np.random.shuffle(inds)) + the env-major -> buffer index map of sf01
 * (ppo2/runner.py:69-74).  out[i] = offset of sample pi(i), pi = keyed Feistel bijection of [0, n) with cycle walking;
 * T > 0: flat index j = e*T + t -> t*N + e (T*N == n); T == 0: out[i] = pi(i). */
int b200rl_shuffle_indices(long long* out, long long n, unsigned long long key, long long T, long long N, void* stream);
int b200rl_counter_add(unsigned long long* ctr, unsigned long long inc, void* stream);

/* per-minibatch advantage moments: ppo2/model.py:136-139.  out = {mean, std} (float64). */
int b200rl_adv_stats(const float* returns, const float* values, const long long* src_idx, long long M, double* out,
                     void* stream);

/* PPO2 loss + gradient w.r.t. head outputs: ppo2/model.py:57-91.  stats[5] += per-sample sums of
 * {pg_loss, vf_loss, entropy, approxkl, clipfrac} (model.py:115); gradients in "sum" scaling. */
int b200rl_cat_loss(const float* logits, long long ld, int nA, const float* vpred, long long ldv,
                    const long long* actions, const long long* src_idx, const float* returns,
                    const float* old_values, const float* old_neglogp, const double* adv_stats, float cliprange,
                    float ent_coef, float vf_coef, void* dlogits, long long ld_dl, void* dv, long long ld_dv,
                    double* stats, long long B, const float* cliprange_dev, void* stream);
int b200rl_gauss_loss(const float* mean, long long ld, const float* logstd, int d, const float* vpred,
                      long long ldv, const float* actions, const long long* src_idx, const float* returns,
                      const float* old_values, const float* old_neglogp, const double* adv_stats, float cliprange,
                      float ent_coef, float vf_coef, void* dmean, long long ld_dm, void* dv, long long ld_dv,
                      float* dlogstd, float inv_M, double* stats, long long B, const float* cliprange_dev,
                      void* stream);

/* optimiser: tf.clip_by_global_norm ppo2/model.py:105-107, tf.clip_by_norm deepq/build_graph.py:416-421,
 * tf.train.AdamOptimizer ppo2/model.py:100 == common/mpi_adam.py:37-42. lr_t = lr*sqrt(1-b2^t)/(1-b1^t). */
int b200rl_sumsq(const float* g, long long n, double* out, void* stream);
int b200rl_seg_sumsq(const float* g, const long long* seg_off, int nseg, double* out, void* stream);
int b200rl_clip_adam(float* p, const float* g, float* m, float* v, long long n, float lr_t, float beta1, float beta2,
                     float eps, float clip, const double* sumsq, const long long* seg_off, int nseg,
                     const float* lr_t_dev, void* stream);
/* acc += g * clip/max(||g||, clip) * weight (clip <= 0: no clipping): the clipped per-microbatch gradients that
 * ppo2/microbatched_model.py:60-70 sums and averages before one apply_gradients. */
int b200rl_clip_accumulate(const float* g, float* acc, long long n, float clip, float weight, const double* sumsq,
                           void* stream);
int b200rl_cast_transpose(const float* src, int R, int C, void* dst, long long ld_dst, void* dstT, long long ld_t,
                          float scale, void* stream);
/* njobs cast_transpose operations in one launch.  jobs: device array of 56-byte records
 *   { const float* src; __half* dst; __half* dstT; long long ld_dst, ld_t; int R, C; float scale; int pad; }
 * (dst / dstT may be NULL); max_rows / max_cols: the largest R / C in the table. */
int b200rl_cast_transpose_batch(const void* jobs, int njobs, int max_rows, int max_cols, void* stream);
int b200rl_cast_f32_f16(const float* src, void* dst, long long rows, int cols, long long ld_src, long long ld_dst,
                        float scale, void* stream);

/* Vector-observation encoding: common/input.py:43-63 (Box -> to_float, Discrete -> one_hot), the optional
 * clip((x - mean) / std, lo, hi) of common/policies.py:182-185, and the minibatch row gather of ppo2/ppo2.py:165.
 * x: float32 [*, raw_dim]; out: fp16 [B, 2*in_pad] = [hi | lo] with hi = fp16(v), lo = fp16(v - hi), so the first
 * GEMM (K = 2*in_pad against [W ; W]) sees the float32 observation to 2^-22 instead of an fp16-rounded copy.
 * onehot_n > 0: x holds the Discrete value (raw_dim = 1), out row = one_hot(x, n). */
int b200rl_obs_encode(const float* x, const long long* src_idx, long long B, int raw_dim, int in_dim, int in_pad,
                      const float* mean, const float* inv_std, float clip_lo, float clip_hi, int onehot_n, void* out,
                      void* stream);

/* prioritized replay: common/segment_tree.py:76-86 (__setitem__), :51-74 (reduce), :105-131
 * (find_prefixsum_idx); deepq/replay_buffer.py:107-115 (_sample_proportional), :157-165 (weights),
 * :169-191 (update_priorities). */
int b200rl_tree_set(double* sum_tree, double* min_tree, long long capacity, const long long* idx, const double* vals,
                    int n, void* stream);
int b200rl_tree_range_sum(const double* tree, long long capacity, long long start, long long end, double* out,
                          void* stream);
int b200rl_per_sample(const double* sum_tree, const double* min_tree, long long capacity, long long n_stored,
                      const double* uniforms, int batch, double beta, long long* idx_out, double* w_out,
                      float* w_out_f32, void* stream);
int b200rl_per_priorities(const float* td, int n, double eps, double alpha, double* powered, double* max_priority,
                          void* stream);

/* DQN: deepq/build_graph.py:388-413 (double-Q target, Huber tf_util.py:39-45, importance weights),
 * deepq/models.py:38-40 (dueling), build_graph.py:184-191 (epsilon-greedy). s_* == NULL -> no dueling. */
int b200rl_dqn_td(const float* a_t, long long lda_t, const float* s_t, long long lds_t, const float* a_on,
                  long long lda_on, const float* s_on, long long lds_on, const float* a_tg, long long lda_tg,
                  const float* s_tg, long long lds_tg, int nA, const long long* idx, const long long* actions,
                  const float* rewards, const float* dones, const float* weights, float gamma, int double_q,
                  float* td_out, void* d_a, long long ld_da, void* d_s, long long ld_ds, double* loss_sum, int B,
                  void* stream);
int b200rl_dqn_act(const float* a, long long lda, const float* s, long long lds, int nA, float eps,
                   unsigned long long seed, unsigned long long step, const float* eps_dev,
                   const unsigned long long* step_dev, long long* actions, int B, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200RL_H */
