#!/usr/bin/env python
"""bench.py -- PPO update env-steps/s (+ GAE steps/s) of the PPO-Dash hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2]

One "step" = one pass of the hot path over one synthetic rollout: compute_returns (GAE) followed
by PPO.update (all epochs x minibatches: gathers, network fwd/bwd, fused loss, clip+Adam).
Workload at N=1: BASELINE.json configs[1], "PPO-Dash full" (recurrent GRU + vector obs, 3x84x84
obs, 32 envs x 512 steps, 8 epochs x 8 minibatches).  N>1: every rank owns the same number of
envs (weak scaling, envs sharded over GPUs), gradients all-reduced with NCCL once per minibatch.

Prints ONE JSON line (rank 0).  `value` = env-steps/s with the rollout resident in HBM; `e2e` = the
same through the public API with the rollout in pinned HOST memory (H2D of every rollout field and
D2H of the losses inside the timed region).  `--impl reference` times the reference algorithm's
CPU path (the oracle port, torch-CPU with all host threads) on a bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c3_12", "c5"])
    ap.add_argument("--precision", default="tf32x3", choices=["fp32", "tf32x3", "tf32"],
                    help="GEMM arithmetic of the policy network (see PolicyEngine.set_precision)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-micro", action="store_true", help="skip the GAE / gather microbenchmarks")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained"), how="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, how="fallback")


class Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------- reference arm (CPU)
def cpu_reference(cfg, minibatches, warm, reps=1, threads=None, device="cpu"):
    """Times the oracle port (torch restatement of PKG/algo/ppo.py + PKG/storage.py) on the box's
    host cores (device="cpu"), or -- as a second comparison row, BASELINE.md section 2 item 5 -- the same
    reference algorithm through STOCK PyTorch CUDA ops (cuDNN conv / GRU, cuBLAS, autograd) on the B200
    (device="cuda").  One sample = `minibatches` minibatches of the update (of epochs*num_mini_batch) plus
    the full compute_returns; the update time is extrapolated linearly to the full update."""
    import numpy as np
    import torch
    from oracle import policy as o_pol
    from oracle import ppo_update as o_upd
    from oracle import returns as o_ret
    from ppodash_b200 import synthetic
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    roll = synthetic.make_rollout(cfg, seed=1234)
    torch.manual_seed(0)
    p = o_pol.init_params(cfg.channels, cfg.num_actions, cfg.vector_obs_len, cfg.recurrent, cfg.hidden_size,
                          concat_vector=cfg.recurrent)
    on_gpu = device != "cpu"
    if on_gpu:
        p = {k: v.to(device) for k, v in p.items()}
        roll_dev = {k: v.to(device) for k, v in roll.items()}
    state = o_upd.UpdateState(p, lr=cfg.lr, eps=cfg.eps)
    times_gae, times_mb = [], []
    total_mb = cfg.ppo_epoch * cfg.num_mini_batch
    for it in range(warm + reps):
        t0 = time.perf_counter()
        ret, v = o_ret.returns_recurrence(roll["rewards"].numpy(), roll["value_preds"].numpy(), roll["masks"].numpy(),
                                          roll["bad_masks"].numpy(), roll["next_value"].numpy(), True, cfg.gamma,
                                          cfg.gae_lambda, False)
        t1 = time.perf_counter()
        r2 = dict(roll_dev if on_gpu else roll)
        r2["returns"] = torch.from_numpy(ret).to(device)
        r2["value_preds"] = torch.from_numpy(v).to(device)
        if on_gpu:
            torch.cuda.synchronize()
            t1 = time.perf_counter()
        torch.manual_seed(99)
        o_upd.ppo_update(state, r2, recurrent=cfg.recurrent, clip_param=cfg.clip_param, ppo_epoch=cfg.ppo_epoch,
                         num_mini_batch=cfg.num_mini_batch, value_loss_coef=cfg.value_loss_coef,
                         entropy_coef=cfg.entropy_coef, max_grad_norm=cfg.max_grad_norm,
                         concat_vector=cfg.recurrent, max_minibatches=minibatches)
        if on_gpu:
            torch.cuda.synchronize()
        t2 = time.perf_counter()
        times_gae.append(t1 - t0)
        times_mb.append((t2 - t1) / minibatches)
    t_gae = min(times_gae)
    t_mb = min(times_mb[warm:]) if warm < len(times_mb) else times_mb[-1]
    t_step = t_gae + t_mb * total_mb
    steps = cfg.num_envs * cfg.num_steps
    return dict(value=steps / t_step, unit="env-steps/s", cores=cores, kind="port",
                sample=f"{minibatches} of {total_mb} minibatches of one {cfg.name} update (+ full compute_returns), "
                       f"best of {len(times_mb) - warm} after {warm} warm-up, extrapolated linearly; "
                       + ("reference algorithm on stock PyTorch CUDA ops (cuDNN/cuBLAS/autograd), compute_returns on host"
                          if on_gpu else "torch-CPU oracle port"),
                gae_steps_per_sec=steps / t_gae, sec_per_minibatch=t_mb, sec_per_step_extrapolated=t_step)


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    # each "step" = a bounded sample (1 minibatch + GAE) of the workload, see cpu_reference()
    res = cpu_reference(cfg, minibatches=1, warm=min(1, args.warmup), reps=max(1, min(args.steps, 3)))
    v = res["value"]
    steps = cfg.num_envs * cfg.num_steps
    line = dict(impl="reference", metric="ppo_update_env_steps_per_sec", value=v, unit="env-steps/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * steps / v, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=cfg.name, envs=cfg.num_envs, num_steps=cfg.num_steps, ppo_epoch=cfg.ppo_epoch,
                            num_mini_batch=cfg.num_mini_batch, recurrent=cfg.recurrent),
                cpu_baseline=dict(value=v, unit="env-steps/s", cores=res["cores"], kind="port", sample=res["sample"]),
                e2e=dict(value=v, unit="env-steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gae_steps_per_sec=res["gae_steps_per_sec"], wall_s=time.perf_counter() - t0)
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- B200 arm
def run_b200(args, cfg):
    import torch
    import torch.distributed as dist
    import ppodash_b200 as ppd
    from ppodash_b200 import _lib, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()

    T, N = cfg.num_steps, cfg.num_envs
    Hs = cfg.hidden_size if cfg.recurrent else 1
    st = ppd.RolloutStorage(T, N, (cfg.channels, cfg.obs_hw, cfg.obs_hw), [cfg.vector_obs_len], Discrete(cfg.num_actions), Hs)
    obs_bytes = (T + 1) * N * cfg.channels * cfg.obs_hw * cfg.obs_hw * 4
    big = obs_bytes > (8 << 30)        # e.g. c5 (42.5 GiB per GPU): generate on the device, no host copy, no e2e leg
    if big:
        st.obs = torch.empty(0)
        st.to(dev)
        gen = torch.Generator(device=dev).manual_seed(1234 + rank)
        st.obs = torch.empty(T + 1, N, cfg.channels, cfg.obs_hw, cfg.obs_hw, device=dev)
        for t0 in range(0, T + 1, 32):
            st.obs[t0:t0 + 32].normal_(generator=gen)
        small = synthetic.make_rollout(cfg, seed=1234 + rank, with_obs=False)
        for k in ppd.RolloutStorage._FIELDS:
            if k != "obs":
                getattr(st, k).copy_(small[k])
        host, nv_host = {}, small["next_value"].pin_memory()
    else:
        roll = synthetic.make_rollout(cfg, seed=1234 + rank)
        host = {k: roll[k].pin_memory() for k in ppd.RolloutStorage._FIELDS}
        nv_host = roll["next_value"].pin_memory()
        st.to(dev)
    torch.manual_seed(0)                      # identical initial weights on every rank
    pol = ppd.Policy((cfg.channels, cfg.obs_hw, cfg.obs_hw), Discrete(cfg.num_actions),
                     base_kwargs={"recurrent": cfg.recurrent, "hidden_size": cfg.hidden_size},
                     vector_obs_len=cfg.vector_obs_len).to(dev)
    pol.engine(args.precision)
    agent = ppd.algo.PPO(pol, cfg.clip_param, cfg.ppo_epoch, cfg.num_mini_batch, cfg.value_loss_coef, cfg.entropy_coef,
                         lr=cfg.lr, eps=cfg.eps, max_grad_norm=cfg.max_grad_norm)
    nv_dev = torch.empty(N, 1, device=dev)
    h2d_bytes = sum(t.numel() * t.element_size() for t in host.values()) + nv_host.numel() * 4

    def upload():
        # public API: small fields at once, observations per env on a copy stream in the order the first epoch consumes them
        st.upload_from(host)
        nv_dev.copy_(nv_host, non_blocking=True)

    def step(with_upload):
        if with_upload:
            upload()
        st.compute_returns(nv_dev, True, cfg.gamma, cfg.gae_lambda, False)
        torch.manual_seed(99)
        return agent.update(st)               # ends with the 3-float device->host read

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, with_upload):
        barrier()
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(nsteps):
            out = step(with_upload)
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), out

    if not big:
        upload()
    else:
        nv_dev.copy_(nv_host)
    for _ in range(args.warmup):
        step(False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.reset_launch_count()
    ms_total, losses = timed(args.steps, False)
    launches = _lib.launch_count()
    ms_e2e = ms_total if big else timed(args.steps, True)[0]
    clocks = sampler.stop() if rank == 0 else None

    # ---- attribute the step to kernels (one extra, untimed-for-the-metric step with per-call CUDA events)
    # (weight-gradient kernels normally run on a side stream; for attribution everything is put on one stream, so a kernel's
    # time is its own and not the time it spent queued behind a concurrent persistent kernel)
    eng = pol.engine()
    ov, eng.overlap_wgrad = eng.overlap_wgrad, False
    with _lib.profiled() as prof:
        step(False)
    eng.overlap_wgrad = ov
    per_kernel = prof.summary()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    steps_per_update = N * T * world
    ms_step = ms_total / args.steps
    value = steps_per_update / (ms_step * 1e-3)
    e2e_value = steps_per_update / (ms_e2e / args.steps * 1e-3)
    total_ms = sum(d["ms"] for d in per_kernel.values())
    shares = sorted(((d["ms"], k, d["calls"]) for k, d in per_kernel.items()), reverse=True)
    top_ms, top_name, top_calls = shares[0]
    # every convolution and every large GEMM of the network is one device kernel (tca_gemm_kernel, csrc/tca_gemm.cu) behind
    # several C-ABI entry points: pool them, they are "the dominant kernel"
    TCA = ("ppd_conv_fwd_nchw", "ppd_conv_fwd_nhwc", "ppd_conv_dgrad_nhwc", "ppd_conv_wgrad", "ppd_tc_gemm", "ppd_tc_gemm_bsplit")
    tca_ms = sum(per_kernel[k]["ms"] for k in TCA if k in per_kernel)
    tca_calls = sum(per_kernel[k]["calls"] for k in TCA if k in per_kernel)
    if args.precision == "tf32x3" and tca_ms >= top_ms:
        top_ms, top_name, top_calls = tca_ms, "tca_gemm_kernel", tca_calls

    # ---- roofline of the dominant kernel (algorithmic work per launch / measured launch time)
    H, V, A, C = cfg.hidden_size, cfg.vector_obs_len, cfg.num_actions, cfg.channels
    E = N // cfg.num_mini_batch
    rows_mb = T * E if cfg.recurrent else (T * N) // cfg.num_mini_batch
    avg_ms = top_ms / top_calls
    roof = dict(kernel=top_name, launches_per_step=top_calls, avg_ms_per_launch=avg_ms,
                share_of_step=top_ms / total_ms if total_ms else None, traffic=None)
    if top_name in ("ppd_gru_forward", "ppd_gru_backward"):
        # recurrent matvec: fwd 2*3H*H flops per (step, env); bwd the same contraction transposed
        flops = 2.0 * 3 * H * H * rows_mb
        ach = flops / (avg_ms * 1e-3) / 1e12
        roof.update(bound="tensor", achieved=ach, peak=pk["bf16"], unit="TFLOP/s", frac=ach / pk["bf16"],
                    latency_us_per_timestep=avg_ms * 1e3 / T,
                    note=f"strictly sequential over T={T} with E={E} envs per minibatch: latency-bound (one grid "
                         f"barrier per timestep), the tensor roofline is not reachable at this E (SURVEY.md 7); "
                         f"peak = bf16 burst {pk['how']}")
    elif top_name == "tca_gemm_kernel":
        # Implicit-GEMM convolutions + the FC / GRU-projection GEMMs, 3xTF32.  Algorithmic bytes: every
        # activation / gradient tensor read once and written once per product, weights once; NO im2col matrices (they do not
        # exist any more).  Pooled over all launches of one step against the measured copy bandwidth.
        Bm = rows_mb
        s1, s2, s3 = 20, 9, 7
        obs_b, a1_b, a2_b, a3_b = Bm * C * cfg.obs_hw ** 2, Bm * s1 * s1 * 32, Bm * s2 * s2 * 64, Bm * s3 * s3 * 32
        K1, K2, K3, FD = C * 64, 512, 576, 1568
        Ip = (H + V + 3) // 4 * 4
        w = dict(c1=32 * K1, c2=64 * K2, c3=32 * K3, fc=H * FD, ih=3 * H * Ip, hh=3 * H * H)
        rec = cfg.recurrent
        fwd_b = (obs_b + w["c1"] + a1_b) + (a1_b + w["c2"] + a2_b) + (a2_b + w["c3"] + a3_b) + (a3_b + w["fc"] + Bm * H) + \
                ((Bm * Ip + w["ih"] + Bm * 3 * H) if rec else 0)
        dgrad_b = (Bm * H + w["fc"] + 2 * a3_b) + (a3_b + w["c3"] + 2 * a2_b) + (a2_b + w["c2"] + 2 * a1_b) + \
                  ((Bm * 3 * H + w["ih"] + Bm * H + Bm * Ip) if rec else 0)          # dx written + ReLU mask read
        wgrad_b = (obs_b + a1_b + w["c1"]) + (a1_b + a2_b + w["c2"]) + (a2_b + a3_b + w["c3"]) + (Bm * H + a3_b + w["fc"]) + \
                  ((Bm * 3 * H + Bm * Ip + w["ih"] + Bm * 3 * H + 2 * Bm * H + w["hh"]) if rec else 0)
        bytes_step = 4.0 * (fwd_b + dgrad_b + wgrad_b) * cfg.ppo_epoch * cfg.num_mini_batch
        ach = bytes_step / (top_ms * 1e-3) / 1e9
        fwd = 2.0 * (819200 * C + 2654208 + 903168 + 802816 + ((3 * H * (H + V)) if rec else 0))
        tf = 3.0 * fwd * Bm * cfg.ppo_epoch * cfg.num_mini_batch / (top_ms * 1e-3) / 1e12
        try:        # DRAM bytes of the same kernel from the committed ncu pass (profiles/), scaled to the launches of one step
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1f_tca_traffic.json")))
            roof["traffic"] = tr["avg_dram_bytes_per_launch"] * top_calls
            roof["traffic_note"] = (f"dram__bytes_read+write.sum averaged over {tr['launches']} consecutive launches (ncu, profiles/"
                                    f"r1f_tca_traffic.md) x {top_calls} launches per step; algorithmic bytes per step {bytes_step:.4g}")
        except Exception:
            pass
        roof.update(bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"], tflops=tf,
                    entry_points={k: round(per_kernel[k]["ms"], 3) for k in TCA if k in per_kernel},
                    note=f"tca_gemm_kernel = persistent tcgen05 3xTF32 kernel (A operand through tensor memory) behind all "
                         f"convolutions (implicit GEMM: TMA im2col views, no im2col matrix in HBM) and the FC / GRU-projection "
                         f"GEMMs, all launches of one step pooled; achieved = algorithmic activation+gradient+weight bytes / summed "
                         f"launch time; peak = copy bandwidth {pk['how']}; useful math {tf:.1f} TFLOP/s fp32-equivalent (x3 on the "
                         f"TF32 pipe).  The conv2-class products (Cout = 64) are nearer the TF32 pipe than HBM: see profiles/.")
    elif top_name in ("ppd_sgemm", "ppd_tc_gemm"):
        # The network's GEMMs are skinny (N = 32 / 64 output channels, or a 2048-row minibatch): ~15-60 FLOP per byte,
        # far left of the tensor ridge (~170 FLOP/B for TF32), i.e. HBM-bound.  Algorithmic bytes = every operand read
        # once and every result written once, fp32, summed over the GEMMs of one minibatch.
        Bm = rows_mb
        s1, s2, s3 = 20, 9, 7
        M1, M2, M3 = Bm * s1 * s1, Bm * s2 * s2, Bm * s3 * s3
        K1, K2, K3, FD = C * 64, 512, 576, 1568
        Ip = (H + V + 3) // 4 * 4
        w = dict(c1=32 * K1, c2=64 * K2, c3=32 * K3, fc=H * FD, ih=3 * H * Ip, hh=3 * H * H)
        fwd_b = (M1 * K1 + w["c1"] + M1 * 32) + (M2 * K2 + w["c2"] + M2 * 64) + (M3 * K3 + w["c3"] + M3 * 32) + \
                (Bm * FD + w["fc"] + Bm * H) + ((Bm * Ip + w["ih"] + Bm * 3 * H) if cfg.recurrent else 0)
        dgrad_b = (Bm * H + w["fc"] + Bm * FD) + (M3 * 32 + w["c3"] + M2 * 64) + (M2 * 64 + w["c2"] + M1 * 32) + \
                  ((Bm * 3 * H + w["ih"] + Bm * H) if cfg.recurrent else 0)
        wgrad_b = (M1 * 32 + M1 * K1 + w["c1"]) + (M2 * 64 + M2 * K2 + w["c2"]) + (M3 * 32 + M3 * K3 + w["c3"]) + \
                  (Bm * H + Bm * FD + w["fc"]) + ((Bm * 3 * H + Bm * Ip + w["ih"] + Bm * 3 * H + 2 * Bm * H + w["hh"])
                                                  if cfg.recurrent else 0)
        bytes_step = 4.0 * (fwd_b + dgrad_b + wgrad_b) * cfg.ppo_epoch * cfg.num_mini_batch
        ach = bytes_step / (top_ms * 1e-3) / 1e9
        fwd = 2.0 * (819200 * C + 2654208 + 903168 + 802816 + ((3 * H * (H + V)) if cfg.recurrent else 0))
        tf = 3.0 * fwd * Bm * cfg.ppo_epoch * cfg.num_mini_batch / (top_ms * 1e-3) / 1e12
        roof.update(bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"], tflops=tf,
                    note=f"GEMM family ({args.precision}; tcgen05 kind::tf32), all launches of one step pooled: skinny GEMMs "
                         f"(N=32/64, or 2048-row minibatch) are HBM-bound; achieved = algorithmic operand+result bytes / summed "
                         f"launch time (side-stream launches overlap, so the sum over-counts time); peak = copy bandwidth "
                         f"{pk['how']}; tensor-side: {tf:.1f} TFLOP/s useful")
    else:
        row_bytes = C * cfg.obs_hw ** 2 * 4 + V * 4 + 8 + 5 * 4
        bytes_launch = 2.0 * row_bytes * rows_mb
        ach = bytes_launch / (avg_ms * 1e-3) / 1e9
        roof.update(bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"],
                    note=f"peak = copy bandwidth {pk['how']}")

    line = dict(metric="ppo_update_env_steps_per_sec", value=value, unit="env-steps/s", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype={"fp32": "f32", "tf32x3": "f32 (3xTF32 tensor-core split, fp32-level accuracy)",
                       "tf32": "tf32"}[args.precision], data="synthetic",
                config=dict(workload=cfg.name, envs_per_gpu=N, num_steps=T, obs=[C, cfg.obs_hw, cfg.obs_hw],
                            vector_obs=V, actions=A, recurrent=cfg.recurrent, ppo_epoch=cfg.ppo_epoch,
                            num_mini_batch=cfg.num_mini_batch, precision=pol.engine().precision,
                            parallelism=f"env-sharded dp{world}",
                            l2="rollout (%.2f GiB) larger than L2; no flush needed" % (h2d_bytes / 2**30)),
                e2e=(dict(value=None, unit="env-steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=12,
                          note="rollout generated on the device (too large for a pinned host copy): no end-to-end leg")
                     if big else
                     dict(value=e2e_value, unit="env-steps/s", h2d_bytes_per_step=h2d_bytes, d2h_bytes_per_step=12,
                          ms_per_step=ms_e2e / args.steps)),
                gpu_launches=launches, roofline=roof, clocks=clocks, losses=list(losses),
                sample_passes_per_sec=value * cfg.ppo_epoch,
                kernel_ms_per_step={k: round(ms, 3) for ms, k, _ in shares})

    # ---- HBM-bound kernels at BASELINE config 4 size (GAE 4096 x 2048) and a gather sweep
    if not args.no_micro:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import microbench as mb
        g = mb.bench_gae(2048, 4096)
        line["gae_steps_per_sec"] = g["steps_per_s"]
        ks = [dict(kernel="returns_scan", config="4096 envs x 2048 steps", bound="hbm", achieved=g["gbs"], peak=pk["hbm"],
                   unit="GB/s", frac=g["gbs"] / pk["hbm"], ms=g["ms"], ms_single_flushed=g["ms_single_flushed"], timing=g["timing"])]
        gg = mb.bench_gather(512, 256, 3, 15, True, 8)
        ks.append(dict(kernel="gather_recurrent", config="256 envs x 512 steps, 3x84x84", bound="hbm", achieved=gg["gbs"],
                       peak=pk["hbm"], unit="GB/s", frac=gg["gbs"] / pk["hbm"], ms_epoch=gg["ms_epoch"]))
        ad = mb.bench_adam(pol.engine().n_params)
        ks.append(dict(kernel="clip_adam", config=f"{pol.engine().n_params} params", bound="hbm", achieved=ad["gbs"],
                       peak=pk["hbm"], unit="GB/s", frac=ad["gbs"] / pk["hbm"], ms=ad["ms"], ms_single_flushed=ad["ms_single_flushed"],
                       timing=ad["timing"]))
        line["kernels"] = ks

    if not args.no_cpu_baseline and world == 1 and not big:
        line["cpu_baseline"] = cpu_reference(cfg, minibatches=2, warm=1)
        # the reference's own setting is ONE thread (run.py:55 torch.set_num_threads(1)); SURVEY.md 8d asks for both
        one = cpu_reference(cfg, minibatches=1, warm=0, threads=1)
        line["cpu_baseline"]["single_thread"] = dict(value=one["value"], unit="env-steps/s", cores=1, sample=one["sample"],
                                                     sec_per_minibatch=one["sec_per_minibatch"])
        torch.set_num_threads(os.cpu_count() or 1)
        # second comparison row: the reference algorithm on the SAME GPU through stock PyTorch (not an optimisation
        # target either; it separates "GPU vs CPU" from "hand-written sm_100a kernels vs stock PyTorch")
        del st, host
        torch.cuda.empty_cache()
        try:
            sc = cpu_reference(cfg, minibatches=4, warm=1, device=str(dev))
            line["stock_torch_cuda_baseline"] = dict(value=sc["value"], unit="env-steps/s", sample=sc["sample"],
                                                     sec_per_minibatch=sc["sec_per_minibatch"])
        except Exception as e:          # a comparison row must never take the bench line down
            line["stock_torch_cuda_baseline"] = dict(value=None, error=str(e)[:200])
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    from ppodash_b200 import synthetic
    cfg = synthetic.CONFIGS[args.workload]
    if args.gpus > 1 and "RANK" not in os.environ and args.impl == "b200":
        port = os.environ.get("MASTER_PORT", "29531")
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port", port,
                                   os.path.abspath(__file__)] + sys.argv[1:])
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_b200(args, cfg)


if __name__ == "__main__":
    main()
