#!/usr/bin/env python
"""Benchmark of the createdb hot path: ProstT5 amino-acid -> 3Di residues/s on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-cpu-baseline] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one pass of the path over the workload's sequences.  Headline workload: BASELINE config 2 (256 synthetic
sequences x 350 aa = 89,600 residues per GPU, one 90,112-token batch); N > 1 = weak scaling, every rank predicts its
count-shard of 256*N such sequences.
  value   device work with tokens and batch tables already in HBM (p5_stage + p5_run_staged, CUDA events on the
          library's launch stream, max over ranks);
  e2e     the public host-buffer call every rank makes (p5_predict_sharded: shard by count, tokenise, pinned H2D,
          forward, D2H and - for N > 1 - the path's single NCCL all-gather of the 3Di bytes, done by the library
          itself, not by torch);
  extras  one timed end-to-end step each of BASELINE config 4 (12,500 ragged 64..1024-aa sequences per GPU = exactly
          config 4 at N = 8) and config 5 (500 sequences of 2000..4000 aa per GPU = config 5 at N = 8), with the
          per-rank device-time spread (load imbalance of count-sharding) and the whole-step tensor-roofline fraction.
Before the line is printed the 3Di letters of the step are compared with the committed oracle fixture
(tests/golden/oracle_letters_full.npz): a step that computes garbage is not timed.
torch is used for the rendezvous, barriers and the max-over-ranks reduction of the timings only.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from unicore_b200 import prostt5_spec as spec, synth  # noqa: E402

METRIC = "ProstT5 3Di residues/sec"
SEQS_PER_GPU, SEQ_LEN = 256, 350
CONFIG4_PER_GPU, CONFIG5_PER_GPU = 12500, 500
MODEL_DIR = os.environ.get("P5_FULL_MODEL_DIR", "/tmp/p5_full_seed1")
LETTERS_FIXTURE = os.path.join(ROOT, "tests", "golden", "oracle_letters_full.npz")
LETTER_MARGIN = 0.05  # = tests/test_gpu_model.py LETTER_MARGIN (twice the full-size logit tolerance)
WORKLOAD = (f"config2: {SEQS_PER_GPU} seqs x {SEQ_LEN} aa per GPU, synthetic ProstT5-shaped weights "
            "(24 layers, d 1024, 32 heads, d_ff 16384), random init seed 1")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"tflops": float(p.get("bf16_tflops_sustained") or p["bf16_tflops"]), "source": "measured (sustained)",
                "hbm_gbs": float(p["hbm_gbs"])}
    return {"tflops": 1400.0, "source": "fallback (sustained)", "hbm_gbs": 6650.0}


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every 100 ms while the timed region runs."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self.err = [], set(), None, None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, repr(e)
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                return
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "error": self.err}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def visible_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except (ValueError, IndexError):
            pass
    return local_rank


def workload(world):
    return spec.synthetic_proteome("config2", n=SEQS_PER_GPU * world)


def check_letters(letters: np.ndarray) -> dict:
    """The first 256 sequences of the step are BASELINE config 2 for every N (the generator is prefix-stable): their
    letters must equal the committed oracle letters wherever the oracle's top-2 margin exceeds LETTER_MARGIN."""
    f = np.load(LETTERS_FIXTURE)
    want, margin = f["config2_letters"], f["config2_margin"]
    got = np.asarray(letters[:len(want)])
    assert len(got) == len(want), "letter count"
    assert set(np.unique(letters)) <= set(b"ACDEFGHIKLMNPQRSTVWY"), "letters outside the 3Di alphabet"
    decided = margin > LETTER_MARGIN
    bad = got != want
    res = {"fixture": "tests/golden/oracle_letters_full.npz (config 2, all 256 sequences)", "residues": int(len(want)),
           "margin": LETTER_MARGIN, "under_margin": int((~decided).sum()),
           "mismatch_under_margin": int((bad & ~decided).sum()), "mismatch_above_margin": int((bad & decided).sum()),
           "sha256": hashlib.sha256(got.tobytes()).hexdigest()}
    if res["mismatch_above_margin"]:
        raise SystemExit("bench.py: the step's 3Di letters differ from the oracle fixture above the margin: %r" % res)
    return res


def c_oracle(threads=None):
    """The C/OpenMP restatement of the path (oracle/prostt5_oracle.c) with every host thread, whatever the launcher
    put into OMP_NUM_THREADS (torchrun sets it to 1)."""
    from oracle import prostt5_oracle_c as OC
    return OC.load_gguf_model(os.path.join(MODEL_DIR, spec.WEIGHT_FILE), threads=threads or os.cpu_count())


def cpu_oracle_rate(n_seqs, seqs_aa, seqs_off):
    """residues/s of the C oracle on the first n_seqs sequences, one sequence at a time as Foldseek's CPU path does."""
    oc = c_oracle()
    oc.predict(seqs_aa[:64].tobytes())  # warm-up (thread pool, page-in of the packed weights)
    t0 = time.perf_counter()
    res = 0
    for i in range(n_seqs):
        s = seqs_aa[int(seqs_off[i]):int(seqs_off[i + 1])].tobytes()
        oc.predict(s)
        res += len(s)
    dt = time.perf_counter() - t0
    return res / dt, res, dt, oc.threads, oc.uses_avx512


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path.  The reference spawns `foldseek createdb --prostt5-model`
    [REF src/modules/createdb.rs:158-166]; neither foldseek nor a Rust toolchain exists here, so this arm times the
    oracle port (oracle/prostt5_oracle.c: fp16 weights x fp32 accumulate, AVX-512, OpenMP on every host core), one
    sequence at a time as Foldseek does, on the sequences of the same config-2 workload: one sequence per step."""
    if rank != 0:
        return
    synth.model_dir(MODEL_DIR, spec.FULL, seed=1)
    oc = c_oracle()
    aa, off = workload(1)
    seq = lambda i: aa[int(off[i]):int(off[i + 1])].tobytes()  # noqa: E731
    for i in range(max(args.warmup, 1)):
        oc.predict(seq(i % SEQS_PER_GPU))
    t0 = time.perf_counter()
    res = 0
    for i in range(args.steps):
        s = seq((args.warmup + i) % SEQS_PER_GPU)
        oc.predict(s)
        res += len(s)
    dt = time.perf_counter() - t0
    v = res / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "residues/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 weights x f32 accumulate (AVX-512 FMA)" if oc.uses_avx512
            else "f16 weights x f32 accumulate", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": "1 sequence of 350 aa of that workload per step"},
            "cpu_baseline": {"value": v, "unit": "residues/s", "cores": oc.threads, "kind": "port",
                             "gflops": res / SEQ_LEN * spec.FULL.flops_per_seq(SEQ_LEN) / dt / 1e9,
                             "sample": f"{args.steps} steps x 1 sequence x 350 aa after {max(args.warmup, 1)} warm-up, "
                                       "C/OpenMP oracle (oracle/prostt5_oracle.c), one sequence at a time"},
            "e2e": {"value": v, "unit": "residues/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, library chatter) was moved to
    stderr by main()."""
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


_REAL_STDOUT = sys.stdout


def git_head():
    """Commit of the tree (here) or of the build that travelled to the GPU box (unicore_b200/lib/BUILD_COMMIT)."""
    try:
        head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True,
                              timeout=5).stdout.strip()
        if head:
            return head
    except Exception:  # noqa: BLE001
        pass
    try:
        return open(os.path.join(ROOT, "unicore_b200", "lib", "BUILD_COMMIT")).read().strip() or None
    except OSError:
        return None


def main():
    global _REAL_STDOUT
    # keep stdout clean for the JSON line: native libraries (NCCL prints its version) write to fd 1
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config-4 / config-5 steps")
    ap.add_argument("--cpu-sample-seqs", type=int, default=6)
    ap.add_argument("--fuse-norm", action="store_true", help="A/B only: RMSNorm inside the residual-add GEMM epilogues (option fuse_norm = 1)")
    ap.add_argument("--attn-impl", type=int, default=-1,
                    help="A/B only: run on libprostt5_b200_debug.so with its option attn_impl (0, 2, 3); never used by the driver")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
        raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, world))

    import torch
    import torch.distributed as dist
    from unicore_b200 import distributed as D
    from unicore_b200.predictor import Comm, Predictor, comm_unique_id

    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        # torch.distributed = rendezvous, barriers and the max over ranks; the data-path collective is the library's own
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        ids = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = Comm(ids[0], rank, world, local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_floats(x):
        if world == 1:
            return [float(x)]
        out = [None] * world
        dist.all_gather_object(out, float(x))
        return out

    if rank == 0:
        synth.model_dir(MODEL_DIR, spec.FULL, seed=1)
    barrier()
    aa_all, off_all = workload(world)
    lens_all = (off_all[1:] - off_all[:-1]).astype(np.int64)
    idx = D.shard_indices(lens_all, rank, world)
    aa, off = D.take_shard(aa_all, off_all, idx)
    total_res = int(off_all[-1])

    pred = Predictor(MODEL_DIR, devices=[local_rank], debug=args.attn_impl >= 0)
    pred.set_option("profile", 1)
    if args.attn_impl >= 0:
        pred.set_option("attn_impl", args.attn_impl)
    if args.fuse_norm:
        pred.set_option("fuse_norm", 1)
    pred.stage(aa, off)
    for _ in range(args.warmup):
        pred.run_staged(None)
    acc = {k: 0.0 for k in ("device_ms", "gemm_ms", "gemm_flops", "gemm_launches", "launches", "attn_ms", "attn_flops",
                            "norm_ms", "head_ms")}
    barrier()
    with ClockSampler(visible_index(local_rank)) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pred.run_staged(None)
            st = pred.stats()
            for k in acc:
                acc[k] += st[k]
        barrier()
        wall = time.perf_counter() - t0
    dev_ms = max_over_ranks(acc["device_ms"])
    wall_ms = max_over_ranks(wall * 1e3)
    value = total_res * args.steps / (dev_ms * 1e-3)

    # end to end through the public host-buffer API: every rank passes the whole proteome, the library shards by count,
    # predicts and (N > 1) all-gathers the 3Di bytes over NCCL
    full = np.zeros(total_res, np.uint8)
    pred.set_option("profile", 0)
    pred.predict_sharded(comm, aa_all, off_all, out=full)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pred.predict_sharded(comm, aa_all, off_all, out=full)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    est = pred.stats()
    letters = check_letters(full)  # raises if the step's letters differ from the oracle fixture above the margin

    pk = peaks()

    def extra_step(name, per_gpu):
        """One end-to-end step of BASELINE config 4 / 5 at this N: count-sharded, through p5_predict_sharded."""
        a_all, o_all = spec.synthetic_proteome(name, n=per_gpu * world)
        lens = (o_all[1:] - o_all[:-1]).astype(np.int64)
        out = np.zeros(int(o_all[-1]), np.uint8)
        warm_a, warm_o = spec.synthetic_proteome(name, n=8 * world)
        pred.predict_sharded(comm, warm_a, warm_o)  # workspace sized, NCCL slabs allocated
        barrier()
        t0 = time.perf_counter()
        pred.predict_sharded(comm, a_all, o_all, out=out)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        st = pred.stats()
        dev = gather_floats(st["device_ms"])
        flops = float(sum(spec.FULL.flops_per_seq(int(L)) for L in lens))
        assert set(np.unique(out)) <= set(b"ACDEFGHIKLMNPQRSTVWY")
        return {"sequences": int(len(lens)), "residues": int(o_all[-1]), "seconds": dt,
                "residues_per_s": float(o_all[-1]) / dt, "pflop": flops / 1e15,
                "tflops_per_gpu": flops / dt / 1e12 / world, "frac_of_sustained_peak": flops / dt / 1e12 / world / pk["tflops"],
                "device_ms_per_rank": dev, "imbalance": (max(dev) / (sum(dev) / len(dev)) - 1.0) if min(dev) > 0 else None,
                "batches_rank0": int(st["batches"]), "sha256": hashlib.sha256(out.tobytes()).hexdigest(),
                "includes": "shard by count + tokenise + H2D + forward + D2H" + (" + NCCL all-gather" if world > 1 else "")}

    extras = None
    if not args.no_extras:
        extras = {"config4": extra_step("config4", CONFIG4_PER_GPU), "config5": extra_step("config5", CONFIG5_PER_GPU)}

    gemm_tflops = acc["gemm_flops"] / (acc["gemm_ms"] * 1e-3) / 1e12
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = tj.get("gemm_dram_bytes_per_launch")
        traffic_src = {"file": "profiles/ncu_traffic.json", "commit": tj.get("commit"), "when": tj.get("when")}
    line = {
        "metric": METRIC, "value": value, "unit": "residues/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16 operands x f32 accumulate (tcgen05 kind::f16)", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "residues_per_step": total_res, "tokens_per_step_per_gpu": int(est["tokens"]),
                   "sharding": "by sequence count, snake order over length-sorted sequences" if world > 1 else "none",
                   "l2": "no flush needed: one step streams 2.4 GB of weights + 6.5 GB of activations (L2 is 126 MB)",
                   "timing": "CUDA events on the library's launch stream per step, summed, max over ranks"},
        "wall_ms_per_step": wall_ms / args.steps,
        "e2e": {"value": total_res * args.steps / e2e_s, "unit": "residues/s",
                "h2d_bytes_per_step": int(est["h2d_bytes"]) * world, "d2h_bytes_per_step": int(est["d2h_bytes"]) * world,
                "includes": "shard + tokenise + pinned H2D + forward + D2H" +
                            (" + the library's NCCL all-gather of the 3Di bytes" if world > 1 else "")},
        "gpu_launches": int(acc["launches"]) * world,
        "letters_check": letters,
        "comm": {"nccl_version": comm.nccl_version, "ranks": comm.world, "owner": "libprostt5_b200.so (ncclAllGather)"} if comm else None,
        "clocks": clk.summary(),
        "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (QKV/O/FFN-in/FFN-out/conv-tap projections)",
                     "achieved": gemm_tflops, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": gemm_tflops / pk["tflops"],
                     "peak_source": pk["source"], "traffic": traffic, "traffic_source": traffic_src,
                     "flops_per_launch": acc["gemm_flops"] / max(acc["gemm_launches"], 1),
                     "ms_per_launch": acc["gemm_ms"] / max(acc["gemm_launches"], 1),
                     "share_of_step": acc["gemm_ms"] / acc["device_ms"],
                     "whole_step_tflops": sum(spec.FULL.flops_per_seq(int(L)) for L in lens_all[idx]) * args.steps
                                          / (acc["device_ms"] * 1e-3) / 1e12,
                     "attention": {"tflops": acc["attn_flops"] / (acc["attn_ms"] * 1e-3) / 1e12 if acc["attn_ms"] else None,
                                   "share_of_step": acc["attn_ms"] / acc["device_ms"]},
                     "norm_share_of_step": acc["norm_ms"] / acc["device_ms"]},
        "extras": extras,
        "commit": git_head(),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, res, dt, threads, avx = cpu_oracle_rate(args.cpu_sample_seqs, aa_all, off_all)
        line["cpu_baseline"] = {"value": rate, "unit": "residues/s", "cores": threads, "kind": "port",
                                "gflops": args.cpu_sample_seqs * spec.FULL.flops_per_seq(SEQ_LEN) / dt / 1e9,
                                "sample": f"first {args.cpu_sample_seqs} sequences of config 2 ({res} residues, {dt:.1f} s), "
                                          "C/OpenMP oracle (oracle/prostt5_oracle.c, "
                                          + ("AVX-512" if avx else "plain C") + "), one sequence at a time"}
    if rank == 0:
        emit(line)
    pred.close()
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
