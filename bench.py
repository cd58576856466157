#!/usr/bin/env python
"""bench.py -- PPO update env-steps/s (+ GAE steps/s) of the PPO-Dash hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2] [--obs u8|f32]

One "step" = one pass of the hot path over one synthetic rollout: compute_returns (GAE) followed
by PPO.update (all epochs x minibatches: gathers, network fwd/bwd, fused loss, clip+Adam).
Workload at N=1: BASELINE.json configs[1], "PPO-Dash full" (recurrent GRU + vector obs + normalised obs,
3x84x84, 32 envs x 512 steps, 8 epochs x 8 minibatches).  N>1: every rank owns the same number of
envs (weak scaling, envs sharded over GPUs), gradients all-reduced with NCCL once per minibatch.

Prints ONE JSON line (rank 0).  `value` = env-steps/s with the rollout resident in HBM; `e2e` = the
same through the public API with the rollout in pinned HOST memory (H2D of every rollout field and
D2H of the losses inside the timed region).  `--impl reference` times the reference algorithm's
CPU path (the oracle port, torch-CPU with all host threads) on a bounded sample of the same workload
and prints the same `config`.

Observations (`--obs`): the PPO-Dash study feeds the policy uint8 frames normalised with the ObtRetro-v6
mean / std (NormalizeWrapper, sohojoe_wrappers.py:855-885).  `u8` (default for 3-channel workloads) stores
the frames as uint8 and normalises on gather (RolloutStorage(obs_dtype=torch.uint8): bit-identical values,
SURVEY.md 8f-2); `f32` stores the reference's float32 tensor.  Both arms see the same VALUES: the CPU
reference is handed the float32 normalised frames.
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
    ap.add_argument("--obs", default="auto", choices=["auto", "u8", "f32"], help="rollout observation storage (module docstring)")
    ap.add_argument("--no-c5", action="store_true", help="skip the BASELINE config 5 block (1024 envs x 512 steps per GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-micro", action="store_true", help="skip the GAE / gather / Adam microbenchmarks")
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


def obs_mode_of(args, cfg):
    if args.obs != "auto":
        return args.obs
    return "u8" if cfg.channels == 3 else "f32"       # the ObtRetro-v6 mean is an RGB (84, 84, 3) table


def config_dict(cfg, args, world, obs_mode):
    """The workload description BOTH arms print (the driver compares them)."""
    return dict(workload=cfg.name, envs_per_gpu=cfg.num_envs, num_steps=cfg.num_steps, obs=[cfg.channels, cfg.obs_hw, cfg.obs_hw],
                vector_obs=cfg.vector_obs_len, actions=cfg.num_actions, recurrent=cfg.recurrent, ppo_epoch=cfg.ppo_epoch,
                num_mini_batch=cfg.num_mini_batch, precision=args.precision, parallelism=f"env-sharded dp{world}",
                obs_values="uint8 frames normalised with the ObtRetro-v6 mean / std" if obs_mode == "u8" else "N(0,1) float32",
                obs_storage="uint8, normalised on gather (bit-identical to the reference's float32 tensor)" if obs_mode == "u8" else "float32",
                l2="rollout larger than L2 (126 MB); no flush needed")


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


# ----------------------------------------------------------------------------- workload
def obs_table():
    """ObtRetro-v6 mean (float64 [3, 84, 84], CHW) and std, as the reference's NormalizeWrapper loads them (from the golden fixture
    tests/golden/obs_pipeline.npz, generated from the reference's own files)."""
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "obs_pipeline.npz"))
    return np.ascontiguousarray(g["mean"].transpose(2, 0, 1)), float(g["std"])


def make_workload(cfg, seed, obs_mode, need_f32):
    """Seeded synthetic rollout on the host.  u8: random uint8 frames [T+1, N, C, H, W] + (if `need_f32`) their float32 normalised
    values, computed as the reference's wrappers do (float64 subtract / divide, one rounding); f32: N(0,1) observations."""
    import torch
    from ppodash_b200 import synthetic
    work = dict(roll=synthetic.make_rollout(cfg, seed=seed, with_obs=(obs_mode == "f32")), frames=None, mean=None, std=None)
    if obs_mode == "u8":
        mean, std = obs_table()
        gen = torch.Generator().manual_seed(seed + 7)
        frames = torch.randint(0, 256, (cfg.num_steps + 1, cfg.num_envs, cfg.channels, cfg.obs_hw, cfg.obs_hw), generator=gen, dtype=torch.uint8)
        work.update(frames=frames, mean=mean, std=std)
        if need_f32:
            m = torch.from_numpy(mean)
            obs = torch.empty(frames.shape, dtype=torch.float32)
            for t0 in range(0, frames.shape[0], 16):
                obs[t0:t0 + 16] = ((frames[t0:t0 + 16].double() - m) / std).float()
            work["roll"]["obs"] = obs
    return work


# ----------------------------------------------------------------------------- reference arm (CPU) / stock-torch row
def cpu_reference(cfg, roll, minibatches, warm, reps=1, threads=None, device="cpu", capture=None, budget_s=None):
    """Times the oracle port (torch restatement of PKG/algo/ppo.py + PKG/storage.py) on the box's host cores (device="cpu"), or --
    as a second comparison row, BASELINE.md section 2 item 5 -- the same reference algorithm through STOCK PyTorch CUDA ops (cuDNN
    conv / GRU with flat weights, cuBLAS, autograd) on the B200 (device="cuda").  One sample = `minibatches` minibatches of the
    update (of epochs * num_mini_batch) plus the full compute_returns; the update time is extrapolated linearly to the full update.
    `capture`: list that receives (value_loss, action_loss, entropy) of every minibatch of the first repetition (parity check).
    `budget_s`: stop repeating once this much wall time is spent (at least one timed repetition is always made)."""
    import torch
    from oracle import policy as o_pol
    from oracle import ppo_update as o_upd
    from oracle import returns as o_ret
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    p = o_pol.init_params(cfg.channels, cfg.num_actions, cfg.vector_obs_len, cfg.recurrent, cfg.hidden_size,
                          concat_vector=cfg.recurrent)
    on_gpu = device != "cpu"
    if on_gpu:
        p = {k: v.to(device) for k, v in p.items()}
        roll_dev = {k: v.to(device) for k, v in roll.items()}
    times_gae, times_mb = [], []
    total_mb = cfg.ppo_epoch * cfg.num_mini_batch
    t_begin = time.perf_counter()
    for it in range(warm + reps):
        if budget_s is not None and it > warm and time.perf_counter() - t_begin > budget_s:
            break
        # fresh parameters every repetition: each one times (and, for `capture`, reports) the same minibatches
        state = o_upd.UpdateState(p, lr=cfg.lr, eps=cfg.eps, flat_gru=on_gpu)
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
        t1b = time.perf_counter()
        torch.manual_seed(99)
        cb = None
        if capture is not None and it == 0:
            cb = lambda k, info: capture.append((info["value_loss"], info["action_loss"], info["entropy"]))
        o_upd.ppo_update(state, r2, recurrent=cfg.recurrent, clip_param=cfg.clip_param, ppo_epoch=cfg.ppo_epoch,
                         num_mini_batch=cfg.num_mini_batch, value_loss_coef=cfg.value_loss_coef,
                         entropy_coef=cfg.entropy_coef, max_grad_norm=cfg.max_grad_norm,
                         concat_vector=cfg.recurrent, max_minibatches=minibatches, on_minibatch=cb)
        if on_gpu:
            torch.cuda.synchronize()
        t2 = time.perf_counter()
        times_gae.append(t1 - t0)
        times_mb.append((t2 - t1b) / minibatches)
    t_gae = min(times_gae)
    t_mb = min(times_mb[warm:]) if warm < len(times_mb) else times_mb[-1]
    t_step = t_gae + t_mb * total_mb
    steps = cfg.num_envs * cfg.num_steps
    return dict(value=steps / t_step, unit="env-steps/s", cores=cores, kind="port",
                sample=f"{minibatches} of {total_mb} minibatches of one {cfg.name} update (+ full compute_returns), "
                       f"best of {len(times_mb) - warm} after {warm} warm-up, extrapolated linearly; "
                       + ("reference algorithm on stock PyTorch CUDA ops (cuDNN conv + GRU with flat weights, cuBLAS, autograd), compute_returns on host"
                          if on_gpu else "torch-CPU oracle port"),
                gae_steps_per_sec=steps / t_gae, sec_per_minibatch=t_mb, sec_per_step_extrapolated=t_step,
                sec_per_sample=t_gae + t_mb * minibatches, reps_timed=len(times_mb) - warm)


def run_reference(args, cfg):
    """The reference arm: every "step" is a bounded sample -- ONE epoch (num_mini_batch minibatches) + the full compute_returns -- of
    the workload, on all host cores; `ms_per_step` is the measured time of a sample, `value` the env-steps/s of the full update
    extrapolated from it (factor ppo_epoch).  Under torchrun only rank 0 runs; it is a single process whatever N is."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    obs_mode = obs_mode_of(args, cfg)
    work = make_workload(cfg, 1234, obs_mode, need_f32=True)
    mbs = cfg.num_mini_batch
    warm = max(0, args.warmup)
    res = cpu_reference(cfg, work["roll"], minibatches=mbs, warm=warm, reps=max(1, args.steps), budget_s=200.0)
    reps = res["reps_timed"]
    one = cpu_reference(cfg, work["roll"], minibatches=1, warm=0, threads=1)
    v = res["value"]
    line = dict(impl="reference", metric="ppo_update_env_steps_per_sec", value=v, unit="env-steps/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * res["sec_per_sample"], higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=config_dict(cfg, args, max(1, args.gpus), obs_mode),
                reference_arithmetic="f32, torch-CPU (oneDNN / MKL)",
                sample_fraction=1.0 / cfg.ppo_epoch, samples_timed=reps,
                note=("ms_per_step is the measured time of one SAMPLE (one epoch of %d minibatches + compute_returns); value = env-steps/s of "
                      "the full %d-epoch update extrapolated from it (x%d), best sample; %d samples timed after %d warm-up (the run stops after "
                      "200 s of samples); single process on the host cores whatever --gpus says" % (mbs, cfg.ppo_epoch, cfg.ppo_epoch, reps, warm)),
                cpu_baseline=dict(value=v, unit="env-steps/s", cores=res["cores"], kind="port", sample=res["sample"],
                                  single_thread=dict(value=one["value"], unit="env-steps/s", cores=1, sample=one["sample"],
                                                     note="the reference's own setting: run.py:55 torch.set_num_threads(1)")),
                e2e=dict(value=v, unit="env-steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gae_steps_per_sec=res["gae_steps_per_sec"], wall_s=time.perf_counter() - t0)
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- B200 arm
def build_storage(ppd, torch, cfg, work, dev, obs_mode):
    """(storage on the device, dict of pinned host fields, next_value pinned)."""
    T, N = cfg.num_steps, cfg.num_envs
    Hs = cfg.hidden_size if cfg.recurrent else 1
    shape = (cfg.channels, cfg.obs_hw, cfg.obs_hw)
    if obs_mode == "u8":
        st = ppd.RolloutStorage(T, N, shape, [cfg.vector_obs_len], Discrete(cfg.num_actions), Hs, obs_dtype=torch.uint8,
                                obs_mean=work["mean"], obs_std=work["std"])
    else:
        st = ppd.RolloutStorage(T, N, shape, [cfg.vector_obs_len], Discrete(cfg.num_actions), Hs)
    roll = work["roll"]
    host = {}
    for k in ppd.RolloutStorage._FIELDS:
        src = work["frames"] if (k == "obs" and obs_mode == "u8") else roll[k]
        host[k] = src.pin_memory()
    st.to(dev)
    return st, host, roll["next_value"].pin_memory()


def run_c5(ppd, torch, dist, args, world, rank, dev):
    """BASELINE config 5 shard on every rank: 1024 envs x 512 steps, uint8 frames generated on the device (10.4 GiB per GPU; the
    float32 layout would be 42.5 GiB), E = 128 envs per minibatch.  1 warm-up + 2 timed steps; no end-to-end leg."""
    from ppodash_b200 import _lib, synthetic
    cfg = synthetic.CONFIGS["c5"]
    mean, std = obs_table()
    T, N = cfg.num_steps, cfg.num_envs
    st = ppd.RolloutStorage(T, N, (3, 84, 84), [cfg.vector_obs_len], Discrete(cfg.num_actions), cfg.hidden_size, obs_dtype=torch.uint8,
                            obs_mean=mean, obs_std=std)
    st.to(dev)
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    for t0 in range(0, T + 1, 64):
        st.obs[t0:t0 + 64].random_(0, 256, generator=gen)
    small = synthetic.make_rollout(cfg, seed=4321 + rank, with_obs=False)
    for k in ppd.RolloutStorage._FIELDS:
        if k != "obs":
            getattr(st, k).copy_(small[k])
    nv = small["next_value"].to(dev)
    torch.manual_seed(0)
    pol = ppd.Policy((3, 84, 84), Discrete(cfg.num_actions), base_kwargs={"recurrent": True, "hidden_size": cfg.hidden_size},
                     vector_obs_len=cfg.vector_obs_len).to(dev)
    pol.engine(args.precision)
    agent = ppd.algo.PPO(pol, cfg.clip_param, cfg.ppo_epoch, cfg.num_mini_batch, cfg.value_loss_coef, cfg.entropy_coef,
                         lr=cfg.lr, eps=cfg.eps, max_grad_norm=cfg.max_grad_norm)

    def step():
        st.compute_returns(nv, True, cfg.gamma, cfg.gae_lambda, False)
        torch.manual_seed(99)
        return agent.update(st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    step()
    nsteps = 2
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(nsteps):
        losses = step()
    b.record()
    barrier()
    ms = torch.tensor([a.elapsed_time(b)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # attribution on ONE stream (as for the main workload): a weight-gradient kernel timed on the side stream would include the time it
    # spent queued behind the input-gradient kernel it alternates with
    eng = pol.engine()
    ov, eng.overlap_wgrad = eng.overlap_wgrad, False
    pg, agent.prefetch_gather = agent.prefetch_gather, False
    with _lib.profiled() as prof:
        step()
    eng.overlap_wgrad, agent.prefetch_gather = ov, pg
    pk = prof.summary()
    ms_step = ms.item() / nsteps
    del st, pol, agent
    torch.cuda.empty_cache()
    return dict(workload=cfg.name, envs_per_gpu=N, num_steps=T, n_gpus=world, steps=nsteps, warmup=1, ms_per_step=ms_step,
                value=N * T * world / (ms_step * 1e-3), unit="env-steps/s", obs_storage="uint8 (10.4 GiB per GPU), generated on the device",
                losses=list(losses), kernel_ms_per_step={k: round(d["ms"], 2) for k, d in sorted(pk.items(), key=lambda kv: -kv[1]["ms"])[:8]},
                note="BASELINE.json config 5 per-GPU shard (weak scaling: every rank runs 1024 envs); device-resident, timed like `value`")


def tca_roofline(cfg, pk, top_ms, top_calls, rows_mb):
    """Algorithmic bytes of the implicit-GEMM convolutions + FC / GRU-projection GEMMs of one step: every
    activation / gradient tensor read once and written once per product, weights once; NO im2col matrices (they do not exist)."""
    H, V, C = cfg.hidden_size, cfg.vector_obs_len, cfg.channels
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
    roof = dict(bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"], tflops=tf, traffic=None,
                algorithmic_bytes_per_step=bytes_step)
    # DRAM bytes of the same kernel from the newest committed ncu pass (profiles/r*_tca_traffic.json), scaled to the launches of a step
    try:
        import glob
        files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_tca_traffic.json")))
        tr = json.load(open(files[-1]))
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from ncu_summary import kernel_source_sha256
        same = tr.get("kernel_source_sha256") == kernel_source_sha256()
        if same:
            roof["traffic"] = tr["avg_dram_bytes_per_launch"] * top_calls
            roof["traffic_note"] = (f"dram__bytes_read+write.sum averaged over {tr['launches']} consecutive launches of one bench step (ncu, "
                                    f"{os.path.basename(files[-1])}; taken from the kernel sources this run was built from: sha256 matches) x "
                                    f"{top_calls} launches per step")
        else:
            # a capture of other kernel sources says nothing about this build: report no traffic rather than a stale one
            roof["traffic_note"] = (f"STALE: {os.path.basename(files[-1])} was captured from other tca_gemm sources than this tree holds "
                                    f"(kernel_source_sha256 differs); re-run tools/profile_gpu.sh and tools/ncu_summary.py traffic")
    except Exception:
        pass
    return roof, tf


def run_b200(args, cfg):
    import numpy as np
    import torch
    import torch.distributed as dist
    import ppodash_b200 as ppd
    from ppodash_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the gradient all-reduce overlaps the convolution backward, whose persistent kernels leave it PPD_COMM_CTAS SMs (default 8)
        os.environ.setdefault("NCCL_MAX_CTAS", os.environ.get("PPD_COMM_CTAS", "8"))
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()
    obs_mode = obs_mode_of(args, cfg)
    if cfg.name.startswith("c5"):
        blk = run_c5(ppd, torch, dist, args, world, rank, dev)
        if rank == 0:
            blk.update(metric="ppo_update_env_steps_per_sec", higher_is_better=True, scaling="weak", data="synthetic",
                       config=config_dict(cfg, args, world, "u8"))
            print(json.dumps(blk), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    T, N = cfg.num_steps, cfg.num_envs
    do_cpu = not args.no_cpu_baseline and world == 1 and rank == 0
    work = make_workload(cfg, 1234 + rank, obs_mode, need_f32=do_cpu)
    st, host, nv_host = build_storage(ppd, torch, cfg, work, dev, obs_mode)

    def new_policy():
        torch.manual_seed(0)                      # identical initial weights on every rank
        p_ = ppd.Policy((cfg.channels, cfg.obs_hw, cfg.obs_hw), Discrete(cfg.num_actions),
                        base_kwargs={"recurrent": cfg.recurrent, "hidden_size": cfg.hidden_size}, vector_obs_len=cfg.vector_obs_len).to(dev)
        p_.engine(args.precision)
        return p_
    pol = new_policy()
    agent = ppd.algo.PPO(pol, cfg.clip_param, cfg.ppo_epoch, cfg.num_mini_batch, cfg.value_loss_coef, cfg.entropy_coef,
                         lr=cfg.lr, eps=cfg.eps, max_grad_norm=cfg.max_grad_norm)
    nv_dev = torch.empty(N, 1, device=dev)
    h2d_bytes = sum(t.numel() * t.element_size() for t in host.values()) + nv_host.numel() * 4

    def upload():
        # public API: small fields at once, observations per env on a copy stream in the order the first epoch consumes them
        st.upload_from(host)
        nv_dev.copy_(nv_host, non_blocking=True)

    def step(with_upload, ag=None):
        if with_upload:
            upload()
        st.compute_returns(nv_dev, True, cfg.gamma, cfg.gae_lambda, False)
        torch.manual_seed(99)
        return (ag or agent).update(st)               # ends with the 3-float device->host read

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, with_upload):
        barrier()
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(nsteps)]
        a.record()
        for i in range(nsteps):
            out = step(with_upload)
            marks[i].record()                      # (every step already ends with a device->host read: no extra synchronisation)
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        each = [round((a if i == 0 else marks[i - 1]).elapsed_time(marks[i]), 3) for i in range(nsteps)]
        return ms.item(), out, each

    upload()
    st.finish_upload()

    # ---- parity at production size, before anything is timed: the first minibatches of the first update from the initial weights
    # against the oracle running the same minibatches on the CPU (rank 0 at N = 1; the same oracle run is the CPU baseline)
    parity, cpu_line = None, None
    if do_cpu:
        pol0 = new_policy()
        eng0 = pol0.engine()
        ag0 = ppd.algo.PPO(pol0, cfg.clip_param, 1, cfg.num_mini_batch, cfg.value_loss_coef, cfg.entropy_coef,
                           lr=cfg.lr, eps=cfg.eps, max_grad_norm=cfg.max_grad_norm)
        got_mb = []
        step0 = ag0.optimizer.step

        def spy(*a, **kw):
            got_mb.append(eng0.flat_grad[eng0.loss_off:eng0.loss_off + 3].clone())
            return step0(*a, **kw)
        ag0.optimizer.step = spy
        step(False, ag0)
        got_mb = [t.cpu().tolist() for t in got_mb]
        del pol0, ag0, eng0
        want_mb = []
        cpu_line = cpu_reference(cfg, work["roll"], minibatches=2, warm=1, capture=want_mb)
        got2, want2 = np.array(got_mb[:2]), np.array(want_mb[:2])
        err = float(np.max(np.abs(got2 - want2) / (np.abs(want2) + 1e-2)))
        parity = dict(minibatches=2, rows_per_minibatch=(T * N // cfg.num_mini_batch), gpu=got2.tolist(), oracle=want2.tolist(),
                      max_err=err, err_definition="|gpu - oracle| / (|oracle| + 1e-2)", gate=1e-4,
                      what="(value_loss, action_loss, entropy) of the first two minibatches of the first update, CUDA path vs CPU oracle, "
                           "same rollout, same permutation")
        if not err <= 1e-4:
            raise SystemExit("bench.py: production-size parity check failed: %s" % json.dumps(parity))

    for _ in range(args.warmup):
        step(False)
    # 12 500 Python-issued launches per update keep the host within ~20 % of the device time, so any host pause inside a step shows
    # (single steps of 130-200 ms were seen on some boxes of the pool).  One candidate is a full-heap garbage collection -- torch alone
    # leaves a few million tracked objects -- so what exists after the warm-up, which is long-lived, is moved out of the collector's
    # sight, as a training script would do.  A precaution: on a quiet box it changes nothing (24 steps each way: 121.2-121.4 ms,
    # slowest step 123.5 / 123.8).  PPD_GC_FREEZE=0: don't.
    if os.environ.get("PPD_GC_FREEZE", "1") != "0":
        import gc
        gc.collect()
        gc.freeze()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.reset_launch_count()
    ms_total, losses, ms_each = timed(args.steps, False)
    launches = _lib.launch_count()
    host_launches = int(_lib.lib().ppd_launch_count())          # issued from Python one by one (the rest were replayed from CUDA graphs)
    ms_e2e, _, ms_each_e2e = timed(args.steps, True)
    clocks = sampler.stop() if rank == 0 else None

    # ---- attribute the step to kernels (one extra, untimed-for-the-metric step with per-call CUDA events)
    # (weight-gradient kernels normally run on a side stream; for attribution everything is put on one stream, so a kernel's
    # time is its own and not the time it spent queued behind a concurrent persistent kernel)
    eng = pol.engine()
    ov, eng.overlap_wgrad = eng.overlap_wgrad, False
    pg, agent.prefetch_gather = agent.prefetch_gather, False          # (the next minibatch's gather normally runs on a side stream, too)
    with _lib.profiled() as prof:
        step(False)
    eng.overlap_wgrad, agent.prefetch_gather = ov, pg
    per_kernel = prof.summary()
    n_params = eng.n_params

    c5 = None
    if not args.no_c5 and cfg.name.startswith("c2"):
        del st
        torch.cuda.empty_cache()
        try:
            c5 = run_c5(ppd, torch, dist, args, world, rank, dev)
        except Exception as e:          # a secondary block must never take the bench line down
            c5 = dict(error=str(e)[:300])

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

    H, V, C = cfg.hidden_size, cfg.vector_obs_len, cfg.channels
    E = N // cfg.num_mini_batch
    rows_mb = T * E if cfg.recurrent else (T * N) // cfg.num_mini_batch
    avg_ms = top_ms / top_calls
    roof = dict(kernel=top_name, launches_per_step=top_calls, avg_ms_per_launch=avg_ms,
                share_of_step=top_ms / total_ms if total_ms else None, traffic=None)
    if top_name in ("ppd_gru_forward", "ppd_gru_backward"):
        flops = 2.0 * 3 * H * H * rows_mb
        ach = flops / (avg_ms * 1e-3) / 1e12
        roof.update(bound="tensor", achieved=ach, peak=pk["bf16"], unit="TFLOP/s", frac=ach / pk["bf16"],
                    latency_us_per_timestep=avg_ms * 1e3 / T,
                    note=f"strictly sequential over T={T} with E={E} envs per minibatch: latency-bound, the tensor roofline is not "
                         f"reachable at this E (SURVEY.md 7); peak = bf16 burst {pk['how']}")
    elif top_name == "tca_gemm_kernel":
        r2, tf = tca_roofline(cfg, pk, top_ms, top_calls, rows_mb)
        roof.update(r2)
        roof.update(entry_points={k: round(per_kernel[k]["ms"], 3) for k in TCA if k in per_kernel},
                    note=f"tca_gemm_kernel = persistent tcgen05 3xTF32 kernel (A operand through tensor memory) behind all "
                         f"convolutions (implicit GEMM: TMA im2col views, no im2col matrix in HBM) and the FC / GRU-projection "
                         f"GEMMs, all launches of one step pooled; achieved = algorithmic activation+gradient+weight bytes / summed "
                         f"launch time; peak = copy bandwidth {pk['how']}; useful math {tf:.1f} TFLOP/s fp32-equivalent (x3 on the "
                         f"TF32 pipe).  The conv2-class products (Cout = 64) are nearer the TF32 pipe than HBM: see profiles/.")
    else:
        row_bytes = C * cfg.obs_hw ** 2 * 4 + V * 4 + 8 + 5 * 4
        bytes_launch = 2.0 * row_bytes * rows_mb
        ach = bytes_launch / (avg_ms * 1e-3) / 1e9
        roof.update(bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"],
                    note=f"peak = copy bandwidth {pk['how']}")

    line = dict(metric="ppo_update_env_steps_per_sec", value=value, unit="env-steps/s", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms_step, ms_each_step=ms_each, ms_each_step_e2e=ms_each_e2e, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype={"fp32": "f32", "tf32x3": "f32 (3xTF32 tensor-core split, fp32-level accuracy)",
                       "tf32": "tf32"}[args.precision], data="synthetic",
                config=config_dict(cfg, args, world, obs_mode),
                e2e=dict(value=e2e_value, unit="env-steps/s", h2d_bytes_per_step=h2d_bytes, d2h_bytes_per_step=12,
                         ms_per_step=ms_e2e / args.steps, ratio_to_value=e2e_value / value),
                gpu_launches=launches, roofline=roof, clocks=clocks, losses=list(losses),
                sample_passes_per_sec=value * cfg.ppo_epoch,
                kernel_ms_per_step={k: round(ms, 3) for ms, k, _ in shares})
    if cfg.recurrent and "ppd_gru_forward" in per_kernel:
        f, b = per_kernel["ppd_gru_forward"], per_kernel["ppd_gru_backward"]
        line["gru"] = dict(fwd_latency_us_per_timestep=1e3 * f["ms"] / (f["calls"] * T), bwd_latency_us_per_timestep=1e3 * b["ms"] / (b["calls"] * T),
                           ms_per_step=f["ms"] + b["ms"], envs_per_minibatch=E,
                           note="T sequential steps per launch: latency-bound at this E; 16-CTA cluster per env, W_hh in registers")
    mg = getattr(agent, "_graphs", None)
    line["minibatch_graphs"] = dict(
        enabled=bool(agent.use_cuda_graph and mg is not None and mg.disabled is None),
        launches_issued_from_python_per_step=host_launches / args.steps, launches_replayed_per_step=(launches - host_launches) / args.steps,
        captures=mg.captures if mg else 0, capture_failed=mg.disabled if mg else None,
        note="PPD_GRAPH=1 (off by default: measured 0.2-0.7 ms per update slower, DESIGN.md section 7): forward + loss + backward of a "
             "minibatch are replayed CUDA graphs (ppodash_b200/minibatch_graph.py); the gathers, clip + Adam and the returns scan are "
             "launched from Python; gpu_launches counts both kinds")
    if parity is not None:
        line["parity_check"] = parity
    if c5 is not None:
        line["c5"] = c5

    # ---- HBM-bound kernels at BASELINE config 4 size (GAE 4096 x 2048), gathers and the optimiser: the LONE flushed launch is the
    # quoted figure; the back-to-back figure over rotating inputs is given beside it
    if not args.no_micro:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import microbench as mb
        g = mb.bench_gae(2048, 4096)
        line["gae_steps_per_sec"] = 2048 * 4096 / (g["ms_single_flushed"] * 1e-3)
        gb = (16 * 2048 * 4096 + 4 * 4096) / 1e9
        ks = [dict(kernel="returns_scan", config="4096 envs x 2048 steps", bound="hbm", achieved=gb / (g["ms_single_flushed"] * 1e-3),
                   peak=pk["hbm"], unit="GB/s", frac=gb / (g["ms_single_flushed"] * 1e-3) / pk["hbm"], ms=g["ms_single_flushed"],
                   timing="lone launch, L2 flushed before it (median of 10)", ms_back_to_back=g["ms"], frac_back_to_back=g["gbs"] / pk["hbm"],
                   timing_back_to_back=g["timing"])]
        gg = mb.bench_gather(512, 256, 3, 15, True, 8)
        ks.append(dict(kernel="gather_recurrent (float32 storage)", config="256 envs x 512 steps, 3x84x84", bound="hbm", achieved=gg["gbs"],
                       peak=pk["hbm"], unit="GB/s", frac=gg["gbs"] / pk["hbm"], ms_epoch=gg["ms_epoch"]))
        gu = mb.bench_gather_u8(512, 256, 3, 15, 8)
        ks.append(dict(kernel="gather_recurrent (uint8 storage, normalise on gather)", config="256 envs x 512 steps, 3x84x84", bound="hbm",
                       achieved=gu["gbs"], peak=pk["hbm"], unit="GB/s", frac=gu["gbs"] / pk["hbm"], ms_epoch=gu["ms_epoch"],
                       note="algorithmic bytes: 1 B read + 4 B written per observation element"))
        ad = mb.bench_adam(n_params)
        ab = 32.0 * n_params / 1e9
        ks.append(dict(kernel="clip_adam", config=f"{n_params} params", bound="hbm", achieved=ab / (ad["ms_single_flushed"] * 1e-3),
                       peak=pk["hbm"], unit="GB/s", frac=ab / (ad["ms_single_flushed"] * 1e-3) / pk["hbm"], ms=ad["ms_single_flushed"],
                       timing="lone launch, L2 flushed before it (median of 10)", ms_back_to_back=ad["ms"], frac_back_to_back=ad["gbs"] / pk["hbm"],
                       timing_back_to_back=ad["timing"]))
        line["kernels"] = ks

    if do_cpu:
        line["cpu_baseline"] = cpu_line
        # the reference's own setting is ONE thread (run.py:55 torch.set_num_threads(1)); SURVEY.md 8d asks for both
        one = cpu_reference(cfg, work["roll"], minibatches=1, warm=0, threads=1)
        line["cpu_baseline"]["single_thread"] = dict(value=one["value"], unit="env-steps/s", cores=1, sample=one["sample"],
                                                     sec_per_minibatch=one["sec_per_minibatch"])
        torch.set_num_threads(os.cpu_count() or 1)
        # second comparison row: the reference algorithm on the SAME GPU through stock PyTorch (not an optimisation
        # target either; it separates "GPU vs CPU" from "hand-written sm_100a kernels vs stock PyTorch")
        del host
        torch.cuda.empty_cache()
        try:
            sc = cpu_reference(cfg, work["roll"], minibatches=16, warm=1, reps=3, device=str(dev))
            line["stock_torch_cuda_baseline"] = dict(value=sc["value"], unit="env-steps/s", sample=sc["sample"],
                                                     sec_per_minibatch=sc["sec_per_minibatch"], ratio_b200_over_stock=value / sc["value"])
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
