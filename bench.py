#!/usr/bin/env python
"""Benchmark of the PPO2 learner hot path (BASELINE.json metric: PPO2 learner env-steps/sec, 4096 envs,
84x84x4 uint8, 128 steps, NatureCNN; configs[1]).

A "step" is ONE full PPO2 update: T+1 batched policy forwards over N envs, the GAE scan, and
noptepochs x nminibatches fused train steps (gather + forward + loss + backward + clip + Adam).

  python bench.py [--gpus N --steps K --warmup W] [--impl reference]
  torchrun --nproc-per-node N bench.py --gpus N ...       (one rank per GPU, NCCL; env-sharded, weak scaling)

Prints ONE JSON line (rank 0).  `value` times the update with the synthetic env resident in HBM; `e2e` times
the same update through the public learn()-style path with a HOST VecEnv (pinned obs -> H2D every env step,
actions D2H every env step, loss statistics D2H every update).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG2 = dict(name="ppo2 NatureCNN synthetic 84x84x4 uint8, 4096 envs x 128 nsteps (BASELINE configs[1])",
            network="cnn", ob_shape=(84, 84, 4), ob_dtype="uint8", n_actions=6, nenvs=4096, nsteps=128,
            nminibatches=4, noptepochs=4, lr=2.5e-4, cliprange=0.1, ent_coef=0.01, vf_coef=0.5, max_grad_norm=0.5,
            gamma=0.99, lam=0.95)
FLOP_FWD_PER_SAMPLE = 18.693e6            # SURVEY.md 8a/8d (conv 15.47 M + fc1 3.21 M + heads 7 k)


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.samples.append(parts)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        pw = [float(s[2]) for s in self.samples if s[2].replace(".", "").isdigit()]
        reasons = []
        for i, nm in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(s[3 + i].lower().startswith("active") for s in self.samples):
                reasons.append(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ---------------------------------------------------------------------------------------------- reference arm
def cpu_reference_update(cfg, nenvs_sample, threads=None, steps=1, warmup=0):
    """The reference's CPU path for one PPO2 update (TF1 unavailable -> oracle port, torch-CPU fp32), with the
    reference's structure: T+1 batched forwards (runner.py:26-50), numpy GAE (:53-65), host shuffle + minibatch
    loop (ppo2.py:157-166) with per-minibatch normalisation, clip, Adam.  Per-sample work identical to cfg;
    only the env count is reduced (bounded sample)."""
    import torch
    from oracle import nets
    from oracle.gae import gae_reference_order, sf01
    if threads:
        torch.set_num_threads(threads)
    T, n = cfg["nsteps"], nenvs_sample
    rng = np.random.RandomState(0)
    np.random.seed(0)
    params = nets.init_policy_params(cfg["network"], cfg["ob_shape"], "discrete", cfg["n_actions"])
    oracle = nets.PPO2Oracle(params, cfg["network"], cfg["ent_coef"], cfg["vf_coef"], cfg["max_grad_norm"])
    pool = [rng.randint(0, 256, (n,) + tuple(cfg["ob_shape"])).astype(np.uint8) for _ in range(8)]
    rews = rng.randn(64, n).astype(np.float32)
    dones = rng.rand(64, n) < 0.01
    nA = cfg["n_actions"]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        mb_obs, mb_act, mb_val, mb_nlp, mb_done, mb_rew = [], [], [], [], [], []
        d = np.zeros(n, bool)
        for t in range(T):
            obs = pool[t % 8]
            a, v, nlp, _ = oracle.step(obs, rng.rand(n, nA).astype(np.float32) * 0.998 + 0.001)
            mb_obs.append(obs.copy()); mb_act.append(a); mb_val.append(v); mb_nlp.append(nlp); mb_done.append(d)
            d = dones[t % 64]
            mb_rew.append(rews[t % 64])
        last_v = oracle.value(pool[T % 8])
        mb_obs, mb_rew, mb_val = np.asarray(mb_obs), np.asarray(mb_rew, np.float32), np.asarray(mb_val, np.float32)
        adv, ret = gae_reference_order(mb_rew, mb_val, np.asarray(mb_done), last_v, d, cfg["gamma"], cfg["lam"])
        obs_f, ret_f, act_f, val_f, nlp_f = map(sf01, (mb_obs, ret, np.asarray(mb_act), mb_val, np.asarray(mb_nlp, np.float32)))
        nbatch = n * T
        nbt = nbatch // cfg["nminibatches"]
        inds = np.arange(nbatch)
        for _ in range(cfg["noptepochs"]):
            np.random.shuffle(inds)
            for s in range(0, nbatch, nbt):
                mb = inds[s:s + nbt]
                oracle.train(cfg["lr"], cfg["cliprange"], obs_f[mb], ret_f[mb], None, act_f[mb], val_f[mb], nlp_f[mb])
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return n * T, times


def pick_cpu_threads():
    """All the host threads the CPU path can USE: torch intra-op threads beyond the physical cores this
    process may run on only add contention (128 threads were 100x slower than 8 on the first box), so probe a
    small conv workload over candidate counts <= the affinity mask and keep the fastest."""
    import torch
    import torch.nn.functional as F
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cands = sorted({c for c in (4, 8, 16, 32, 64, avail) if c <= avail}) or [1]
    x = torch.randn(64, 4, 84, 84)
    w = torch.randn(32, 4, 8, 8)
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        F.conv2d(x, w, stride=4)
        t0 = time.perf_counter()
        for _ in range(3):
            F.conv2d(x, w, stride=4)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CFG2
    n = args.ref_envs
    threads = pick_cpu_threads()
    nb, times = cpu_reference_update(cfg, n, threads=threads, steps=args.steps, warmup=min(args.warmup, 1))
    ms = 1000.0 * float(np.mean(times))
    val = nb / (ms / 1000.0)
    sample = f"{n} envs x {cfg['nsteps']} steps per update (same per-sample work as 4096 envs), torch-CPU fp32 oracle port"
    out = {"impl": "reference", "metric": "PPO2 learner env-steps/sec", "value": val, "unit": "env-steps/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": ms,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": cfg["name"], "sample": sample},
           "cpu_baseline": {"value": val, "unit": "env-steps/s", "cores": torch.get_num_threads(), "kind": "port",
                            "sample": sample},
           "e2e": {"value": val, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------- our arm
def main():
    # some images export NCCL_DEBUG=VERSION, which prints a banner on stdout next to the JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nenvs", type=int, default=CFG2["nenvs"], help="envs per GPU (default: BASELINE config)")
    ap.add_argument("--ref-envs", type=int, default=16, help="envs in the bounded CPU-reference sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="skip the per-kernel CUDA-event profile")
    ap.add_argument("--no-targets", action="store_true", help="skip the stand-alone GAE / fc1 / PER microbenchmarks")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    from baselines_b200 import _lib
    from baselines_b200.common.policies import build_policy
    from baselines_b200.common.vec_env import DeviceSyntheticVecEnv, SyntheticVecEnv
    from baselines_b200.ppo2.model import Model
    from baselines_b200.ppo2.ppo2 import run_epochs
    from baselines_b200.ppo2.runner import Runner

    cfg = dict(CFG2)
    cfg["nenvs"] = args.nenvs
    N, T = cfg["nenvs"], cfg["nsteps"]
    nbatch = N * T
    nbatch_train = nbatch // cfg["nminibatches"]
    dev = torch.device("cuda", local_rank)
    np.random.seed(0)

    def make(env):
        policy = build_policy(env, cfg["network"])
        model = Model(policy=policy, ob_space=env.observation_space, ac_space=env.action_space, nbatch_act=N,
                      nbatch_train=nbatch_train, nsteps=T, ent_coef=cfg["ent_coef"], vf_coef=cfg["vf_coef"],
                      max_grad_norm=cfg["max_grad_norm"], comm=None if world > 1 else False)
        return model, Runner(env=env, model=model, nsteps=T, gamma=cfg["gamma"], lam=cfg["lam"])

    def update(model, runner):
        ro, _ = runner.run_device()
        st = run_epochs(model, ro, cfg["lr"], cfg["cliprange"], nbatch, nbatch_train, cfg["noptepochs"], dev)
        return torch.stack(st).mean(dim=0)

    def timed(model, runner, steps, warmup, read_back, profile=False):
        for _ in range(warmup):
            update(model, runner)
        torch.cuda.synchronize()
        if profile:
            _lib.profile_begin()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.LAUNCHES
        e0.record()
        for _ in range(steps):
            st = update(model, runner)
            if read_back:
                st.cpu()                                         # the loss statistics a user reads each update
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps, (_lib.LAUNCHES - l0)

    # ---- kernel-only / device-resident value ------------------------------------------------------------
    env_d = DeviceSyntheticVecEnv(N, cfg["ob_shape"], np.uint8, cfg["n_actions"], seed=rank, device=dev)
    model, runner = make(env_d)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    prof = None
    ms_step, launches = timed(model, runner, args.steps, args.warmup, read_back=False, profile=not args.no_profile)
    if not args.no_profile:
        prof = _lib.profile_end()
    clocks = sampler.stop() if rank == 0 else None
    value = world * nbatch / (ms_step / 1000.0)

    # ---- e2e through the host VecEnv ---------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        del runner, env_d
        torch.cuda.empty_cache()
        ob_bytes = int(np.prod(cfg["ob_shape"]))
        common_h2d = T * N * 5 + cfg["noptepochs"] * nbatch * 8          # rewards + dones + minibatch permutations
        # (a) the reference's Atari pipeline (run.py build_env): VecFrameStack(venv, 4).  The env produces ONE new
        #     84x84x1 frame per step; our VecFrameStack keeps the stack in HBM, so only new frames cross PCIe.
        e2e_stacked = None
        if len(cfg["ob_shape"]) == 3 and cfg["ob_shape"][-1] == 4:
            from baselines_b200.common.vec_env import VecFrameStack
            frame_shape = tuple(cfg["ob_shape"][:-1]) + (1,)
            env_f = VecFrameStack(SyntheticVecEnv(N, frame_shape, np.uint8, cfg["n_actions"], seed=rank), 4)
            runner_f = Runner(env=env_f, model=model, nsteps=T, gamma=cfg["gamma"], lam=cfg["lam"])
            ms_f, _ = timed(model, runner_f, max(1, args.steps), 1, read_back=True)
            e2e = {"value": world * nbatch / (ms_f / 1000.0), "unit": "env-steps/s", "ms_per_step": ms_f,
                   "h2d_bytes_per_step": T * N * (ob_bytes // 4 + 1) + common_h2d,
                   "d2h_bytes_per_step": T * N * 8 + 40,
                   "input": "VecFrameStack(host VecEnv of 84x84x1 frames, 4): new frames uploaded, stack kept in HBM"}
            del runner_f, env_f
            torch.cuda.empty_cache()
        # (b) a host VecEnv that hands out full stacked observations: every 84x84x4 observation is uploaded
        env_h = SyntheticVecEnv(N, cfg["ob_shape"], np.uint8, cfg["n_actions"], seed=rank)
        runner_h = Runner(env=env_h, model=model, nsteps=T, gamma=cfg["gamma"], lam=cfg["lam"])
        ms_e2e, _ = timed(model, runner_h, max(1, args.steps), 1, read_back=True)
        e2e_stacked = {"value": world * nbatch / (ms_e2e / 1000.0), "unit": "env-steps/s", "ms_per_step": ms_e2e,
                       "h2d_bytes_per_step": (T + 1) * N * ob_bytes + common_h2d,
                       "d2h_bytes_per_step": T * N * 8 + 40,
                       "input": "host VecEnv handing out stacked observations: all of them uploaded"}
        if e2e is None:
            e2e = e2e_stacked
        else:
            e2e["stacked_upload"] = e2e_stacked

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (CUDA events recorded live in the timed region) ---------------
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks.update(json.load(open(pk)))
        peaks["src"] = "measured"
    roofline, kernels = None, None
    if prof:
        kernels = summarize_profile(prof, args.steps)
        tot = sum(k["ms_per_step"] for k in kernels.values())
        for k in kernels.values():
            k["share"] = k["ms_per_step"] / tot if tot else 0.0
        top = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
        tk = kernels[top]
        # the kernel is judged against the roof it sits closer to: algorithmic flops vs the measured dense-bf16
        # throughput, algorithmic bytes vs the measured HBM copy bandwidth (both fractions are kept)
        n_launch = max(1.0, tk["launches_per_step"])
        f_tensor = f_hbm = 0.0
        if tk.get("flops_per_step"):
            f_tensor = tk["flops_per_step"] / (tk["ms_per_step"] / 1e3) / 1e12 / peaks["bf16_tflops_sustained"]
        if tk.get("bytes_per_step"):
            f_hbm = tk["bytes_per_step"] / (tk["ms_per_step"] / 1e3) / 1e9 / peaks["hbm_gbs"]
        traffic = None
        tj = os.path.join(ROOT, "profiles", "traffic.json")       # ncu --set full dram bytes, per algorithmic byte
        if os.path.exists(tj):
            t = json.load(open(tj)).get(top)
            if t and tk.get("bytes_per_step"):
                traffic = t["dram_bytes_per_algorithmic_byte"] * tk["bytes_per_step"] / n_launch
        if f_tensor >= f_hbm:
            roofline = {"kernel": top, "bound": "tensor", "achieved": f_tensor * peaks["bf16_tflops_sustained"],
                        "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": f_tensor,
                        "peak_src": peaks["src"] + " (sustained cuBLAS bf16)"}
        else:
            roofline = {"kernel": top, "bound": "hbm", "achieved": f_hbm * peaks["hbm_gbs"], "peak": peaks["hbm_gbs"],
                        "unit": "GB/s", "frac": f_hbm, "peak_src": peaks["src"] + " (sustained copy)"}
        roofline.update({"traffic": traffic, "algorithmic_bytes_per_launch": tk.get("bytes_per_step", 0.0) / n_launch,
                         "algorithmic_flops_per_launch": tk.get("flops_per_step", 0.0) / n_launch,
                         "frac_tensor": f_tensor, "frac_hbm": f_hbm, "ms_per_launch": tk["ms_per_step"] / n_launch,
                         "share_of_step": tk["share"]})

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:              # rank 0 at N=1 only (contract)
        n = args.ref_envs
        nb, times = cpu_reference_update(cfg, n, threads=pick_cpu_threads(), steps=1, warmup=0)
        cpu_baseline = {"value": nb / times[0], "unit": "env-steps/s", "cores": torch.get_num_threads(),
                        "kind": "port", "sample": f"one PPO2 update on {n} envs x {T} steps (same per-sample work), "
                                                  f"torch-CPU fp32 oracle port of the TF1 graph"}

    targets = None
    try:
        if args.no_targets or world > 1:
            raise RuntimeError("skipped (--no-targets or N > 1: stand-alone kernels are a 1-GPU measurement)")
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import microbench
        mb = microbench.run(quick=True)
        g3 = mb["gae"][1]
        targets = {"gae_cfg3": {"T": g3["T"], "N": g3["N"], "ms": g3["ms"], "achieved_gbs": g3["gbs"],
                                "peak_gbs": peaks["hbm_gbs"], "frac": g3["gbs"] / peaks["hbm_gbs"], "target": 0.6},
                   "gae_cfg2": {"ms": mb["gae"][0]["ms"], "achieved_gbs": mb["gae"][0]["gbs"],
                                "note": "8.9 MB: L2-resident / launch-bound, not an HBM measurement"},
                   "fc1": [{"kind": c["kind"], "M": c["M"], "tflops": c["tflops"], "peak": peaks["bf16_tflops"],
                            "frac": c["tflops"] / peaks["bf16_tflops"], "target": 0.5} for c in mb["fc1"]],
                   "gae_cpu_numpy": mb.get("gae_cpu"),          # reference numpy loop (oracle port), host, same sizes
                   "per_cfg4": mb.get("per"),                   # PER sample/update at capacity 2^20 vs the python port
                   "replay_gather_cfg4": mb.get("replay_gather"),   # obs gather + cast of one replay sample
                   "dqn_cfg4": mb.get("dqn"),                   # one deepq train step at batch 512
                   "how": mb["l2_flush"] + "; CUDA events per launch, median of 10 after 3 warm-ups"}
    except Exception as ex:                                    # never lose the headline line to an extra
        targets = {"error": repr(ex)}

    out = {"metric": "PPO2 learner env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate (GAE f64 carry)",
           "data": "synthetic",
           "config": {"workload": cfg["name"], "envs_per_gpu": N, "nsteps": T, "nminibatches": cfg["nminibatches"],
                      "noptepochs": cfg["noptepochs"], "parallelism": f"dp{world} (env-sharded, grad allreduce)",
                      "l2": "inputs larger than L2 (rollout obs 14.8 GB, every minibatch streams 3.7 GB)",
                      "train_chunk": model.chunk},
           "tflops_per_step": 127.5 * N / 4096, "gpu_launches": launches, "clocks": clocks, "e2e": e2e,
           "roofline": roofline, "cpu_baseline": cpu_baseline, "targets": targets, "kernels": kernels}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def summarize_profile(prof, steps):
    """prof: {label: [ms, calls, flops, bytes]} accumulated over the timed region."""
    out = {}
    for label, (ms, calls, flops, nbytes) in prof.items():
        out[label] = {"ms_per_step": ms / steps, "launches_per_step": calls / steps,
                      "flops_per_step": flops / steps, "bytes_per_step": nbytes / steps}
        if flops:
            out[label]["tflops"] = flops / (ms / 1e3) / 1e12 if ms else None
        if nbytes:
            out[label]["gbs"] = nbytes / (ms / 1e3) / 1e9 if ms else None
    return out


if __name__ == "__main__":
    main()
