#!/usr/bin/env python
"""bench.py -- the headline measurement of the B200 statevector path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the CPU reference arm

Workload (BASELINE.json configs[1]): GHZ preparation + QFT on 30 qubits, complex64, one
B200; with N > 1 GPUs the register grows by log2(N) qubits (weak scaling: 2**30 amplitudes
per GPU), the state is sharded by its high positions and global qubits are swapped in with
an NCCL all-to-all.  A "step" simulates the whole circuit on a fresh |0...0> state.

One JSON line on stdout (rank 0).  `value` is device-timed with the compiled passes already
resident in HBM; `e2e` goes through the public API with host inputs (circuit object in,
sampled counts dict out) and includes planning, the H2D copy of the pass descriptors, the
sweep, device-side sampling and the D2H copy of the samples.
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

import numpy as np

METRIC = "statevector_gates_per_s"
UNIT = "gates/s"
LOCAL_QUBITS = 30
CPU_SAMPLE_QUBITS = int(os.environ.get("QB2_CPU_SAMPLE_QUBITS", 26))  # (tests shrink it)
E2E_SHOTS = 4096


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks during the timed region (recipe: B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self, first=0, last=None):
        """Summary of the samples [first, last) (the timed region); when the region was too short to
        catch one, of every sample since the warm-up began (same workload), and the note says so."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = self.lines[first:last] if last is not None else self.lines[first:]
        note = "samples inside the timed region"
        if not lines:
            lines = self.lines
            note = "timed region shorter than one sampling period: samples since the start of the warm-up (same workload)"
        for line in lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(smax)) if smax else None,
            "power_w_max": float(max(power)) if power else None,
            "samples": len(sm),
            "note": note,
            "reasons": sorted(reasons),
        }


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------
# CPU arm: the numpy oracle (port of the reference's algorithm) on the host cores
# ----------------------------------------------------------------------------------------


def cpu_gates_per_s(sample_qubits, budget_s=25.0):
    """Times the oracle's restatement of the reference's fast path
    (oracle.ref_sim.statevector_reference_path: grouped unitaries, sparse contraction while the
    state is sparse, dense afterwards) on the whole GHZ+QFT circuit at `sample_qubits`, and
    scales gates/s to the 30-qubit workload by amplitude count (every dense contraction is
    linear in the state size: tensor_factor.py:67-89).  `budget_s` picks the sample width: the
    widest register whose predicted time (4x per two qubits) stays inside the budget."""
    from oracle import ref_sim
    from qrisp_b200.circuit import ghz_qft_circuit

    qc = ghz_qft_circuit(sample_qubits)
    n_gates = sum(1 for i in qc.data if i.matrix is not None)
    t0 = time.perf_counter()
    ref_sim.statevector_reference_path(qc)
    dt = time.perf_counter() - t0
    return n_gates / dt, n_gates, dt


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info

        n = max((p.get("num_threads", 1) for p in threadpool_info()), default=1)
        return int(n)
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    n_total = LOCAL_QUBITS + int(np.log2(args.gpus))
    vals = []
    # the sample register shrinks with the number of steps so that the whole run stays within a few
    # minutes (about 5-10 s per step at 26 qubits, a factor 2 per qubit)
    sample_qubits = CPU_SAMPLE_QUBITS if args.steps <= 10 else (CPU_SAMPLE_QUBITS - 1 if args.steps <= 25 else CPU_SAMPLE_QUBITS - 2 if args.steps <= 60 else CPU_SAMPLE_QUBITS - 4)
    for _ in range(min(1, args.warmup)):
        cpu_gates_per_s(sample_qubits - 4)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        g, done, dt = cpu_gates_per_s(sample_qubits)
        vals.append(g)
    wall = time.perf_counter() - t_all
    scale = 2.0 ** (sample_qubits - n_total)
    value = float(np.mean(vals)) * scale
    sample = (f"oracle/ref_sim.statevector_reference_path (numpy port of the reference's grouped, sparse-then-dense "
              f"contraction) on the whole GHZ+QFT circuit at {sample_qubits} qubits, gates/s scaled by "
              f"2**({sample_qubits}-{n_total}) to the {n_total}-qubit workload (dense contractions are linear in "
              f"the state size)")
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / max(1, args.steps),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "complex64",
        "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    n_total = LOCAL_QUBITS + int(np.log2(n_gpus))
    return {
        "workload": f"GHZ + QFT on {n_total} qubits, dense statevector, complex64 (BASELINE.json configs[1]"
        + ("" if n_gpus == 1 else f", widened by log2({n_gpus}) qubits for weak scaling") + ")",
        "qubits": n_total,
        "local_qubits": LOCAL_QUBITS,
        "state_bytes_per_gpu": 8 << LOCAL_QUBITS,
        "l2": "state (8 GiB per GPU) is 65x the 126 MB L2: no flush needed between steps",
        "parallelism": "single GPU" if n_gpus == 1 else f"state sharded by the top {int(np.log2(n_gpus))} positions",
    }


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------


def run_gpu_arm(args):
    import torch

    from qrisp_b200 import _lib
    from qrisp_b200.circuit import ghz_qft_circuit

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        # stdout carries exactly one JSON line: whatever libraries print there (NCCL's version banner
        # when the image sets NCCL_DEBUG) goes to stderr; the line is written to the saved descriptor
        sys.stdout.flush()
        _real_stdout = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        sys.stdout = _real_stdout

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    lib = _lib.load()
    n_total = LOCAL_QUBITS + int(np.log2(world))
    qc = ghz_qft_circuit(n_total)
    instrs = [i for i in qc.data if i.matrix is not None]
    n_gates = len(instrs)

    if world == 1:
        from qrisp_b200.engine import StateVector

        sv = StateVector(n_total)
    else:
        from qrisp_b200.distributed import ShardedStateVector

        sv = ShardedStateVector(n_total, group=dist.group.WORLD)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- compile once: the timed region replays the resident program
    prog = sv.compile(instrs)

    def step():
        sv.reset()
        prog.run()

    # the clock sampler streams from before the warm-up (nvidia-smi takes a moment to start); the
    # samples that fall inside the timed region are picked out by line count below
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    lib.qb2_reset_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 2)]
    pass_ms = []
    barrier()
    ev_start, ev_end = ev[-2], ev[-1]
    mark0 = len(sampler.lines)
    ev_start.record()
    marks = []
    for k in range(args.steps):
        sv.reset()
        ev[2 * k].record()
        prog.run()
        ev[2 * k + 1].record()
    ev_end.record()
    barrier()
    launches = int(lib.qb2_launch_count())
    mark1 = len(sampler.lines)
    clocks = sampler.stop(mark0, mark1) if rank == 0 else None
    total_ms = ev_start.elapsed_time(ev_end)
    sweep_ms = sum(ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(args.steps))
    t = torch.tensor([total_ms, sweep_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, sweep_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / args.steps
    value = n_gates / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (pass_kernel): algorithmic bytes = read + write of
    # the local state once per launch
    n_pass = prog.n_passes
    state_bytes = 8 << LOCAL_QUBITS
    comm_ms = None
    if world > 1:
        # one extra, untimed replay with events around every qubit swap
        sv.reset()
        prog.run(time_comm=True)
        c = torch.tensor([prog.comm_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.MAX)
        comm_ms = float(c[0])
    pass_avg_ms = (sweep_ms / args.steps - (comm_ms or 0.0)) / max(1, n_pass)
    peak, peak_src = measured_peaks()
    # algorithmic bytes: every pass reads and writes the local state once, except the first pass of a
    # step, which synthesises |0...0> instead of reading it (QB2_FEATURE_ZERO_INPUT)
    first_reads = 0
    bytes_per_launch = (2 * n_pass - 1 + first_reads) * state_bytes / max(1, n_pass)
    achieved = bytes_per_launch / (pass_avg_ms * 1e-3) / 1e9
    # one more, untimed replay with events around every pass: where the step's time goes
    per_pass_ms = None
    if world == 1:
        sv.reset()
        prog._enter()
        pev = []
        for i in range(len(prog.steps)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s_, o_, h_ = prog.steps[i], prog.offsets[i], prog._headers[i]
            if s_.kind == "pass":
                sv._run_pass(s_, h_, prog.dev, o_, sparse=prog._entry_sparse if i == 0 else None)
            else:
                sv._run_big(s_)
            e1.record()
            pev.append((e0, e1))
        torch.cuda.synchronize()
        sv.pos[:] = prog.pos_out
        per_pass_ms = [round(a.elapsed_time(b), 3) for a, b in pev]
    roofline = {
        "kernel": "pass_kernel<float,RB=5,256,2,lean>: packed FP32x2 arithmetic, 2 CTAs x 256 threads per SM, 32 amplitudes per thread",
        "bound": "hbm",
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": None,
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": bytes_per_launch,
        "algorithmic_bytes_note": "2 x local state per pass; the first pass of a step synthesises |0...0> instead of reading it",
        "avg_launch_ms": pass_avg_ms,
        "per_pass_ms": per_pass_ms,
        "passes_per_step": n_pass,
        "kernel_ops_per_step": prog.n_kernel_ops,
        "phases_per_step": prog.n_phases,
    }
    traffic_path = os.path.join(ROOT, "profiles", "pass_kernel_traffic.json")
    if os.path.exists(traffic_path):
        try:
            roofline["traffic"] = json.load(open(traffic_path)).get("dram_bytes_per_launch")
        except Exception:
            pass

    # ---- end to end through the public API: circuit object in, counts dict out
    from qrisp_b200.circuit import Circuit

    e2e = None
    if world == 1:
        from qrisp_b200.simulator import sample_counts

        qc_m = Circuit(n_total)
        qc_m.data = list(qc.data)
        qc_m.measure(list(range(n_total)))
        del sv, prog
        torch.cuda.empty_cache()
        sample_counts(qc_m, 16, seed=0)  # warm-up
        torch.cuda.synchronize()
        reps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for i in range(reps):
            counts = sample_counts(qc_m, E2E_SHOTS, seed=i)
            assert sum(counts.values()) == E2E_SHOTS
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / reps
        from qrisp_b200 import simulator as _s

        e2e = {
            "value": n_gates / e2e_s,
            "unit": UNIT,
            "ms_per_step": 1e3 * e2e_s,
            "h2d_bytes_per_step": int(_s.LAST_H2D_BYTES),
            "d2h_bytes_per_step": int(_s.LAST_D2H_BYTES),
            "api": f"qrisp_b200.simulator.sample_counts(circuit, shots={E2E_SHOTS}): plan + upload + sweep + device sampling",
        }
    else:
        from qrisp_b200.distributed import sample_counts_sharded

        qc_m = Circuit(n_total)
        qc_m.data = list(qc.data)
        qc_m.measure(list(range(n_total)))
        n_swaps = prog.n_swaps
        del sv, prog
        torch.cuda.empty_cache()
        counts, sv2 = sample_counts_sharded(qc_m, 16, seed=0, group=dist.group.WORLD)  # warm-up (NCCL buffers, allocator)
        del sv2
        barrier()
        t0 = time.perf_counter()
        counts, sv2 = sample_counts_sharded(qc_m, E2E_SHOTS, seed=0, group=dist.group.WORLD)
        barrier()
        e2e_s = time.perf_counter() - t0
        assert sum(counts.values()) == E2E_SHOTS
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {
            "value": n_gates / float(t[0]),
            "unit": UNIT,
            "ms_per_step": 1e3 * float(t[0]),
            "h2d_bytes_per_step": int(sv2.io["h2d"]),
            "d2h_bytes_per_step": int(sv2.io["d2h"]),
            "api": f"qrisp_b200.distributed.sample_counts_sharded(circuit, shots={E2E_SHOTS}) on every rank, after one warm-up call",
            "qubit_swaps": n_swaps,
        }
        del sv2

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- CPU baseline on this box's host cores (bounded sample)
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        g, done, dt = cpu_gates_per_s(CPU_SAMPLE_QUBITS)
        scale = 2.0 ** (CPU_SAMPLE_QUBITS - n_total)
        cpu_baseline = {
            "value": g * scale,
            "unit": UNIT,
            "cores": cpu_threads(),
            "kind": "port",
            "sample": f"oracle/ref_sim.statevector_reference_path (grouped, sparse-then-dense) on all {done} gates of "
            f"GHZ+QFT at {CPU_SAMPLE_QUBITS} qubits ({dt:.1f} s), gates/s scaled by 2**({CPU_SAMPLE_QUBITS}-{n_total}) "
            f"to the {n_total}-qubit workload",
        }

    line = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "complex64",
        "data": "synthetic",
        "config": workload_config(world),
        "gates_per_step": n_gates,
        "amp_updates_per_s": value * float(1 << n_total),
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "e2e": e2e,
    }
    if comm_ms is not None:
        line["comm_ms_per_step"] = comm_ms
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.gpus < 1 or args.gpus & (args.gpus - 1):
        raise SystemExit("--gpus must be a power of two")
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
