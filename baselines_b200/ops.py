"""Torch-tensor front end of the C-ABI: tensors only supply device pointers and the current stream.

Every function launches hand-written sm_100a kernels from libb200rl.so; nothing here computes with
torch ops (torch is plumbing: allocation, streams, torch.distributed).
"""
import torch

from . import _lib

MODE_F16_ACT, MODE_F32_STORE, MODE_F32_ATOMIC, MODE_F16_DACT, MODE_F16_SHUFFLE = 0, 1, 2, 3, 4
ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2
ACT_CODES = {None: ACT_NONE, "none": ACT_NONE, "relu": ACT_RELU, "tanh": ACT_TANH}


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (the hot path has no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected {dtype}, got {t.dtype}")


def gae_scan(rewards, values, dones, last_values, last_dones, advs, returns, gamma, lam, variant=-1):
    """rewards/values/advs/returns float32 [T,N]; dones uint8 [T,N]; last_* [N]."""
    T, N = rewards.shape
    for t, dt, nm in ((rewards, torch.float32, "rewards"), (values, torch.float32, "values"),
                      (dones, torch.uint8, "dones"), (last_values, torch.float32, "last_values"),
                      (last_dones, torch.uint8, "last_dones"), (advs, torch.float32, "advs"),
                      (returns, torch.float32, "returns")):
        _chk(t, dt, nm)
        assert t.is_contiguous(), nm
    _lib.call("b200rl_gae_scan", _ptr(rewards), _ptr(values), _ptr(dones), _ptr(last_values), _ptr(last_dones),
              _ptr(advs), _ptr(returns), T, N, float(gamma), float(lam), int(variant), _stream(),
              label="gae_scan", nbytes=17.0 * T * N + 5.0 * N)


def gemm(A, B, C, *, M, N, K, lda, ldb, ldc, bias=None, saved=None, ld_saved=0, mn_major=False,
         mode=MODE_F16_ACT, act=ACT_NONE, alpha=1.0, split_k=1, max_ctas=0, tag=None, remap=(0, 0, 0), saved_bits=None):
    _chk(A, torch.float16, "A")
    _chk(B, torch.float16, "B")
    _chk(bias, torch.float32, "bias")
    _chk(saved, torch.float16, "saved")
    _chk(saved_bits, torch.int16, "saved_bits")
    _lib.call("b200rl_gemm_f16", _ptr(A), _ptr(B), _ptr(C), _ptr(bias), _ptr(saved), int(M), int(N), int(K),
              int(lda), int(ldb), int(ldc), int(ld_saved), int(bool(mn_major)), int(mode), int(act), float(alpha),
              int(split_k), int(max_ctas), int(remap[0]), int(remap[1]), int(remap[2]), _ptr(saved_bits), _stream(),
              label="gemm." + (tag or ("wgrad" if mn_major else "tn")),
              flops=2.0 * M * N * K,
              nbytes=2.0 * (M * K + N * K) + M * N * (2 if mode in (MODE_F16_ACT, MODE_F16_DACT) else 4)
              + ((0.125 if saved_bits is not None else 2.0) * M * N if mode == MODE_F16_DACT else 0))


def conv_gemm(x, B, H, W, C, R, S, stride_h, stride_w, pad_h, pad_w, OH, OW, wt_or_dz, ldb, out, ldc, N, kind,
              mode, act=ACT_NONE, alpha=1.0, bias=None, saved=None, ld_saved=0, split_k=1, shuffle=None, tag=None):
    """Implicit-GEMM convolution (TMA im2col A operand).  kind 0: forward / data-gradient form; kind 1: wgrad."""
    _chk(x, torch.float16, "x")
    _chk(wt_or_dz, torch.float16, "wt_or_dz")
    sh = shuffle or (0, 0, 0, 0)
    rows = B * OH * OW
    _lib.call("b200rl_conv_gemm", _ptr(x), int(B), H, W, C, R, S, stride_h, stride_w, pad_h, pad_w, OH, OW,
              _ptr(wt_or_dz), int(ldb), _ptr(out), int(ldc), _ptr(bias), _ptr(saved), int(ld_saved), int(N), int(kind),
              int(mode), int(act), float(alpha), int(split_k), int(sh[0]), int(sh[1]), int(sh[2]), int(sh[3]),
              _stream(), label="conv." + (tag or str(kind)), flops=2.0 * rows * N * R * S * C,
              nbytes=2.0 * B * H * W * C + 2.0 * R * S * C * N + rows * N * (2.0 if kind == 0 else 2.0))


import ctypes as _C


def _iarr(vals, ctype):
    return (ctype * len(vals))(*[int(v) for v in vals])


def _u8_args(u8):
    """u8 = (images uint8 [.., H, W, C], src_idx or None, H, W, C, s) or None."""
    if u8 is None:
        return None, None, 0, 0, 0, 0
    x, idx, H, W, C, s = u8
    _chk(x, torch.uint8, "u8 images")
    _chk(idx, torch.int64, "u8 src_idx")
    return _ptr(x), _ptr(idx), int(H), int(W), int(C), int(s)


def conv_shift_fwd(X, B, Hg, Wg, C, W, ldw, N, shifts, vy, vx, out, omap, *, saved=None, smap=None, bias=None,
                   act=ACT_NONE, dact=False, alpha=1.0, tag=None, u8=None, bits_out=None, saved_bits=None,
                   useful_rows=None, kx=1):
    """Shift-GEMM convolution (forward, or data gradient with dact=True).  omap / smap: 6-tuples
    (mode, sN, sY, sX, Cq, s).  useful_rows (accounting only): positions that are real conv outputs -- the kernel
    also computes (and discards) the grid positions that are not; reported flops / bytes count the useful ones."""
    _chk(X, torch.float16, "X")
    _chk(W, torch.float16, "W")
    _chk(out, torch.float16, "out")
    u8a = _u8_args(u8)
    _chk(bits_out, torch.int16, "bits_out")
    _chk(saved_bits, torch.int16, "saved_bits")
    sh = _iarr(shifts, _C.c_int)
    om = _iarr(omap, _C.c_longlong)
    sm = _iarr(smap, _C.c_longlong) if smap is not None else None
    rows = B * Hg * Wg
    useful = float(useful_rows) if useful_rows is not None else float(B) * vy * vx
    if dact:      # data gradient: reads the valid dY rows, writes every dX position, reads the activation mask
        nbytes = 2.0 * useful * C + 2.0 * rows * N
        nbytes += 0.0 if saved is None and saved_bits is None else (0.125 if saved_bits is not None else 2.0) * rows * N
    else:         # forward: reads every input position, writes the valid outputs (+ 1 bit per element of mask)
        nbytes = (1.0 if u8 is not None else 2.0) * rows * C + 2.0 * useful * N + (0.125 * useful * N if bits_out is not None else 0.0)
    _lib.call("b200rl_conv_shift_fwd", _ptr(X), int(B), Hg, Wg, C, _ptr(W), int(ldw), int(N), len(shifts), sh, vy, vx,
              _ptr(out), om, _ptr(saved), sm, _ptr(bias), int(act), int(bool(dact)), float(alpha), *u8a,
              _ptr(bits_out), _ptr(saved_bits), int(kx), _stream(),
              label="convs." + (tag or "fwd"), flops=2.0 * useful * N * len(shifts) * kx * C, nbytes=nbytes)


def conv_shift_wgrad(X, rows, C, dY, N, shifts, G, ldg, alpha=1.0, max_ctas=0, tag=None, gbias=None, alpha_b=1.0,
                     u8=None, useful_rows=None, kx=1):
    _chk(X, torch.float16, "X")
    _chk(dY, torch.float16, "dY")
    _chk(G, torch.float32, "G")
    sh = _iarr(shifts, _C.c_int)
    _chk(gbias, torch.float32, "gbias")
    u8a = _u8_args(u8)
    _lib.call("b200rl_conv_shift_wgrad", _ptr(X), int(rows), C, _ptr(dY), int(N), len(shifts), sh, _ptr(G), int(ldg),
              float(alpha), _ptr(gbias), float(alpha_b), int(max_ctas), *u8a, int(kx), _stream(),
              label="convs." + (tag or "wgrad"),
              flops=2.0 * (float(useful_rows) if useful_rows is not None else rows) * N * len(shifts) * kx * C,
              nbytes=rows * (1.0 if u8 is not None else 2.0) * C +
              2.0 * N * (float(useful_rows) if useful_rows is not None else rows))


def dgrad_weights(w, out, R, S, Cin, Cout, s, ld):
    _chk(w, torch.float32, "w")
    _chk(out, torch.float16, "out")
    _lib.call("b200rl_dgrad_weights", _ptr(w), _ptr(out), R, S, Cin, Cout, s, int(ld), _stream())


def _conv_out(H, W, rf, stride, same_pad):
    if same_pad:
        return -(-H // stride), -(-W // stride)
    return (H - rf) // stride + 1, (W - rf) // stride + 1


def im2col(x, cols, B, H, W, C, rf, stride, same_pad=False, src_idx=None, tag=None):
    _chk(src_idx, torch.int64, "src_idx")
    _chk(cols, torch.float16, "cols")
    src_u8 = x.dtype == torch.uint8
    if not src_u8:
        _chk(x, torch.float16, "x")
    OH, OW = _conv_out(H, W, rf, stride, same_pad)
    _lib.call("b200rl_im2col", _ptr(x), int(src_u8), _ptr(src_idx), _ptr(cols), int(B), H, W, C, rf, stride,
              int(bool(same_pad)), _stream(), label="im2col." + (tag or ("u8" if src_u8 else "f16")),
              nbytes=float(B) * H * W * C * (1 if src_u8 else 2) + 2.0 * B * OH * OW * rf * rf * C)


def frame_stack(prev, frame, news, out, nstack, c):
    """out = VecFrameStack update of prev with the new frames (vec_frame_stack.py:17-25); uint8 [N, ..., nstack*c]."""
    for t, nm in ((prev, "prev"), (frame, "frame"), (news, "news"), (out, "out")):
        _chk(t, torch.uint8, nm)
    N = out.shape[0]
    pixels = out[0].numel() // (nstack * c)
    if prev.numel() != out.numel() or frame.numel() != N * pixels * c or news.numel() != N:
        raise RuntimeError("frame_stack: shape mismatch")
    _lib.call("b200rl_frame_stack", _ptr(prev), _ptr(frame), _ptr(news), _ptr(out), int(N), int(pixels), int(nstack),
              int(c), _stream(), label="frame_stack", nbytes=float(2 * out.numel() + frame.numel()))


def s2d_gather(x, out, B, H, W, C, s, src_idx=None):
    _chk(x, torch.uint8, "x")
    _chk(out, torch.float16, "out")
    _chk(src_idx, torch.int64, "src_idx")
    _lib.call("b200rl_s2d_gather", _ptr(x), _ptr(src_idx), _ptr(out), int(B), H, W, C, s, _stream(),
              label="s2d_gather", nbytes=3.0 * B * H * W * C)


def col2im(dcols, saved, dx, B, H, W, C, rf, stride, same_pad=False, act=ACT_NONE, tag=None):
    _chk(dcols, torch.float16, "dcols")
    _chk(dx, torch.float16, "dx")
    OH, OW = _conv_out(H, W, rf, stride, same_pad)
    _lib.call("b200rl_col2im", _ptr(dcols), _ptr(saved), _ptr(dx), int(B), H, W, C, rf, stride, int(bool(same_pad)),
              int(act), _stream(), label="col2im." + (tag or ""),
              nbytes=2.0 * B * OH * OW * rf * rf * C + 2.0 * B * H * W * C * (2 if saved is not None else 1))


def colsum(dz, db, rows, C, ld, alpha=1.0):
    _chk(dz, torch.float16, "dz")
    _chk(db, torch.float32, "db")
    _lib.call("b200rl_colsum", _ptr(dz), _ptr(db), int(rows), int(C), int(ld), float(alpha), _stream(),
              label="colsum", nbytes=2.0 * rows * C)


def cat_step(logits, ld, nA, vpred, ldv, actions, values, neglogp, B, uniforms=None, seed=0, offset=0,
             offset_dev=None):
    _chk(logits, torch.float32, "logits")
    _chk(actions, torch.int64, "actions")
    _chk(uniforms, torch.float32, "uniforms")
    _chk(offset_dev, torch.int64, "offset_dev")
    _lib.call("b200rl_cat_step", _ptr(logits), int(ld), int(nA), _ptr(vpred), int(ldv), _ptr(uniforms), int(seed),
              int(offset), _ptr(offset_dev), _ptr(actions), _ptr(values), _ptr(neglogp), int(B), _stream())


def gauss_step(mean, ld, logstd, d, vpred, ldv, actions, values, neglogp, B, normals=None, seed=0, offset=0,
               offset_dev=None):
    _chk(mean, torch.float32, "mean")
    _chk(actions, torch.float32, "actions")
    _chk(normals, torch.float32, "normals")
    _chk(offset_dev, torch.int64, "offset_dev")
    _lib.call("b200rl_gauss_step", _ptr(mean), int(ld), _ptr(logstd), int(d), _ptr(vpred), int(ldv), _ptr(normals),
              int(seed), int(offset), _ptr(offset_dev), _ptr(actions), _ptr(values), _ptr(neglogp), int(B), _stream())


def set_scalars(dst, *vals):
    """dst[0..len(vals)) = vals (float32 device tensor): values travel as kernel arguments."""
    _chk(dst, torch.float32, "dst")
    v = [float(x) for x in vals] + [0.0] * (4 - len(vals))
    _lib.call("b200rl_set_scalars", _ptr(dst), len(vals), v[0], v[1], v[2], v[3], _stream())


def shuffle_indices(out, n, key, T=0, N=0):
    """out[:n] = buffer offsets of a keyed pseudo-random permutation of the n rollout samples (ppo2.py:160)."""
    _chk(out, torch.int64, "out")
    _lib.call("b200rl_shuffle_indices", _ptr(out), int(n), int(key) & 0xFFFFFFFFFFFFFFFF, int(T), int(N), _stream(),
              label="shuffle_indices", nbytes=8.0 * n)


def counter_add(ctr, inc=1):
    _chk(ctr, torch.int64, "ctr")
    _lib.call("b200rl_counter_add", _ptr(ctr), int(inc), _stream())


def adv_stats(returns, values, src_idx, M, out):
    _chk(out, torch.float64, "out")
    _chk(src_idx, torch.int64, "src_idx")
    _lib.call("b200rl_adv_stats", _ptr(returns), _ptr(values), _ptr(src_idx), int(M), _ptr(out), _stream())


def cat_loss(logits, ld, nA, vpred, ldv, actions, src_idx, returns, old_values, old_neglogp, adv_st, cliprange,
             ent_coef, vf_coef, dlogits, ld_dl, dv, ld_dv, stats, B, cliprange_dev=None):
    _chk(actions, torch.int64, "actions")
    _chk(stats, torch.float64, "stats")
    _lib.call("b200rl_cat_loss", _ptr(logits), int(ld), int(nA), _ptr(vpred), int(ldv), _ptr(actions), _ptr(src_idx),
              _ptr(returns), _ptr(old_values), _ptr(old_neglogp), _ptr(adv_st), float(cliprange), float(ent_coef),
              float(vf_coef), _ptr(dlogits), int(ld_dl), _ptr(dv), int(ld_dv), _ptr(stats), int(B), _ptr(cliprange_dev),
              _stream())


def gauss_loss(mean, ld, logstd, d, vpred, ldv, actions, src_idx, returns, old_values, old_neglogp, adv_st,
               cliprange, ent_coef, vf_coef, dmean, ld_dm, dv, ld_dv, dlogstd, inv_M, stats, B, cliprange_dev=None):
    _chk(actions, torch.float32, "actions")
    _lib.call("b200rl_gauss_loss", _ptr(mean), int(ld), _ptr(logstd), int(d), _ptr(vpred), int(ldv), _ptr(actions),
              _ptr(src_idx), _ptr(returns), _ptr(old_values), _ptr(old_neglogp), _ptr(adv_st), float(cliprange),
              float(ent_coef), float(vf_coef), _ptr(dmean), int(ld_dm), _ptr(dv), int(ld_dv), _ptr(dlogstd),
              float(inv_M), _ptr(stats), int(B), _ptr(cliprange_dev), _stream())


def sumsq(g, out):
    _chk(g, torch.float32, "g")
    _chk(out, torch.float64, "out")
    _lib.call("b200rl_sumsq", _ptr(g), g.numel(), _ptr(out), _stream(), label="sumsq", nbytes=4.0 * g.numel())


def seg_sumsq(g, seg_off, nseg, out):
    _chk(seg_off, torch.int64, "seg_off")
    _lib.call("b200rl_seg_sumsq", _ptr(g), _ptr(seg_off), int(nseg), _ptr(out), _stream())


def clip_adam(p, g, m, v, lr_t, beta1, beta2, eps, clip, sumsq_buf, seg_off=None, nseg=0, lr_t_dev=None):
    for t, nm in ((p, "p"), (g, "g"), (m, "m"), (v, "v")):
        _chk(t, torch.float32, nm)
    _lib.call("b200rl_clip_adam", _ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), float(lr_t), float(beta1),
              float(beta2), float(eps), float(clip if clip else 0.0), _ptr(sumsq_buf), _ptr(seg_off), int(nseg),
              _ptr(lr_t_dev), _stream(), label="clip_adam", nbytes=28.0 * p.numel())


def clip_accumulate(g, acc, clip, weight, sumsq_buf):
    _chk(g, torch.float32, "g")
    _chk(acc, torch.float32, "acc")
    _lib.call("b200rl_clip_accumulate", _ptr(g), _ptr(acc), g.numel(), float(clip if clip else 0.0), float(weight),
              _ptr(sumsq_buf), _stream(), label="clip_accumulate", nbytes=12.0 * g.numel())


_cast_recording = None      # list of jobs while a CastPlan is being recorded


def cast_transpose(src, R, C, dst, ld_dst, dstT, ld_t, scale=1.0):
    _chk(src, torch.float32, "src")
    if _cast_recording is not None:
        _cast_recording.append((src, int(R), int(C), dst, int(ld_dst), dstT, int(ld_t), float(scale)))
        return
    _lib.call("b200rl_cast_transpose", _ptr(src), int(R), int(C), _ptr(dst), int(ld_dst), _ptr(dstT), int(ld_t),
              float(scale), _stream())


class CastPlan:
    """The fp16 operand refresh of a whole network (every cast_transpose its layers issue) as ONE launch: the calls are
    recorded once into a device table of jobs (pointers are fixed for the life of the network)."""

    def __init__(self, fn, device):
        global _cast_recording
        import numpy as np
        _cast_recording = []
        try:
            fn()
            jobs = _cast_recording
        finally:
            _cast_recording = None
        self.n = len(jobs)
        self.keep = jobs                                   # keeps the tensors (and their storage) alive
        rec = np.zeros(max(self.n, 1), dtype=np.dtype([("src", "<u8"), ("dst", "<u8"), ("dstT", "<u8"), ("ld_dst", "<i8"),
                                                        ("ld_t", "<i8"), ("R", "<i4"), ("C", "<i4"), ("scale", "<f4"),
                                                        ("pad", "<i4")]))
        assert rec.dtype.itemsize == 56
        for i, (src, Rr, Cc, dst, ld_dst, dstT, ld_t, scale) in enumerate(jobs):
            rec[i] = (src.data_ptr(), 0 if dst is None else dst.data_ptr(), 0 if dstT is None else dstT.data_ptr(),
                      ld_dst, ld_t, Rr, Cc, scale, 0)
        self.table = torch.from_numpy(rec.view(np.uint8).copy()).to(device)
        self.max_r = max([j[1] for j in jobs], default=1)
        self.max_c = max([j[2] for j in jobs], default=1)

    def run(self):
        if self.n:
            _lib.call("b200rl_cast_transpose_batch", _ptr(self.table), self.n, self.max_r, self.max_c, _stream(),
                      label="cast_transpose_batch")


def cast_f32_f16(src, dst, rows, cols, ld_src, ld_dst, scale=1.0):
    _chk(src, torch.float32, "src")
    _chk(dst, torch.float16, "dst")
    _lib.call("b200rl_cast_f32_f16", _ptr(src), _ptr(dst), int(rows), int(cols), int(ld_src), int(ld_dst),
              float(scale), _stream())


def obs_encode(x, out, B, raw_dim, in_dim, in_pad, src_idx=None, mean=None, inv_std=None, clip=(-5.0, 5.0), onehot_n=0):
    """float32 observation rows -> fp16 [hi | lo] operand rows (input.py:43-63, policies.py:182-185, ppo2.py:165)."""
    _chk(x, torch.float32, "x")
    _chk(out, torch.float16, "out")
    _chk(src_idx, torch.int64, "src_idx")
    _chk(mean, torch.float32, "mean")
    _chk(inv_std, torch.float32, "inv_std")
    _lib.call("b200rl_obs_encode", _ptr(x), _ptr(src_idx), int(B), int(raw_dim), int(in_dim), int(in_pad), _ptr(mean),
              _ptr(inv_std), float(clip[0]), float(clip[1]), int(onehot_n), _ptr(out), _stream(),
              label="obs_encode", nbytes=float(B) * (4.0 * raw_dim + 4.0 * in_pad))


def tree_set(sum_tree, min_tree, capacity, idx, vals):
    _chk(sum_tree, torch.float64, "sum_tree")
    _chk(idx, torch.int64, "idx")
    _chk(vals, torch.float64, "vals")
    _lib.call("b200rl_tree_set", _ptr(sum_tree), _ptr(min_tree), int(capacity), _ptr(idx), _ptr(vals), idx.numel(),
              _stream())


def tree_range_sum(tree, capacity, start, end, out):
    _lib.call("b200rl_tree_range_sum", _ptr(tree), int(capacity), int(start), int(end), _ptr(out), _stream())


def per_sample(sum_tree, min_tree, capacity, n_stored, uniforms, beta, idx_out, w_out, w_out_f32=None):
    _chk(uniforms, torch.float64, "uniforms")
    _chk(idx_out, torch.int64, "idx_out")
    _chk(w_out, torch.float64, "w_out")
    _lib.call("b200rl_per_sample", _ptr(sum_tree), _ptr(min_tree), int(capacity), int(n_stored), _ptr(uniforms),
              uniforms.numel(), float(beta), _ptr(idx_out), _ptr(w_out), _ptr(w_out_f32), _stream())


def per_priorities(td, eps, alpha, powered, max_priority):
    _chk(td, torch.float32, "td")
    _lib.call("b200rl_per_priorities", _ptr(td), td.numel(), float(eps), float(alpha), _ptr(powered),
              _ptr(max_priority), _stream())


def dqn_td(a_t, lda_t, s_t, lds_t, a_on, lda_on, s_on, lds_on, a_tg, lda_tg, s_tg, lds_tg, nA, idx, actions,
           rewards, dones, weights, gamma, double_q, td_out, d_a, ld_da, d_s, ld_ds, loss_sum, B):
    _lib.call("b200rl_dqn_td", _ptr(a_t), int(lda_t), _ptr(s_t), int(lds_t), _ptr(a_on), int(lda_on), _ptr(s_on),
              int(lds_on), _ptr(a_tg), int(lda_tg), _ptr(s_tg), int(lds_tg), int(nA), _ptr(idx), _ptr(actions),
              _ptr(rewards), _ptr(dones), _ptr(weights), float(gamma), int(bool(double_q)), _ptr(td_out), _ptr(d_a),
              int(ld_da), _ptr(d_s), int(ld_ds), _ptr(loss_sum), int(B), _stream())


def dqn_act(a, lda, s, lds, nA, eps, seed, step, actions, B, eps_dev=None, step_dev=None):
    _chk(eps_dev, torch.float32, "eps_dev")
    _chk(step_dev, torch.int64, "step_dev")
    _lib.call("b200rl_dqn_act", _ptr(a), int(lda), _ptr(s), int(lds), int(nA), float(eps), int(seed), int(step),
              _ptr(eps_dev), _ptr(step_dev), _ptr(actions), int(B), _stream())
