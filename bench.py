#!/usr/bin/env python3
"""bench.py — mutated genome Gbp/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c1|c5]

A step = one pass of the hot path over the whole synthetic genome: sample ->
resolve -> link -> plan -> splice/emit FASTA image -> format VCF, all on the GPU.
`value`  : input bases / device time with the genome already resident in HBM.
`e2e`    : the same through the C ABI with HOST buffers — pinned H2D of the genome,
           D2H of the FASTA image and the VCF body inside the timed region.
`roofline`: the splice/emit kernel's algorithmic bytes / its CUDA-event time vs the
           measured HBM copy bandwidth (MEASURED_PEAKS.json).
`cpu_baseline`: the C oracle port (oracle/ms_oracle.c) on a bounded sample, 1 thread.
`--impl reference`: the reference's path on the host CPU = the oracle port on all host
           threads (the reference is pure Python and cannot travel to the GPU box).
N > 1 (torchrun): contigs are partitioned over ranks (LPT by length, no data-path
collective); total work is fixed, so scaling is "strong".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

GRCH38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
          133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
          58617616, 64444167, 46709983, 50818468, 156040895, 57227415]
GRCH38_NAMES = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY"]

WORKLOADS = {
    # BASELINE.json configs[1]: ARGS, all mutation types, GRCh38-shaped 3.1 Gbp / 24 contigs
    "c2": dict(desc="C2: ARGS all types (sn .01 titv 2 in/de .001 len<=10, du/iv/tl .0005 len<=50) on synthetic GRCh38-shaped 3.088 Gbp, 24 contigs",
               lengths=GRCH38, names=GRCH38_NAMES, rates6=[0.01, 0.001, 0.001, 0.0005, 0.0005, 0.0005],
               minlen=[1, 1, 1, 2, 1, 1, 1], maxlen=[1, 10, 10, 50, 50, 50, 50], titv=2.0, n_fraction=0.03, telomere=10000),
    # BASELINE.json configs[0]: the reference's CPU-runnable case
    "c1": dict(desc="C1: ARGS sn .01 in/de .001 len 1-10 on a synthetic 10 Mbp contig",
               lengths=[10_000_000], names=["chr1"], rates6=[0.01, 0.001, 0.001, 0.0, 0.0, 0.0],
               minlen=[1, 1, 1, 2, 1, 1, 1], maxlen=[1, 10, 10, 3, 2, 2, 2], titv=1.0, n_fraction=0.0, telomere=0),
    # BASELINE.json configs[4]: many small contigs
    "c5": dict(desc="C5: ARGS all types on 200k contigs x 5 kbp",
               lengths=[5000] * 200_000, names=None, rates6=[0.01, 0.001, 0.001, 0.0005, 0.0005, 0.0005],
               minlen=[1, 1, 1, 2, 1, 1, 1], maxlen=[1, 10, 10, 50, 50, 50, 50], titv=2.0, n_fraction=0.0, telomere=0),
}


def type_cdf(rates6):
    rates = list(rates6[:5]) + [rates6[5] / 2, rates6[5] / 2]     # rmt.py:91-94
    total = sum(rates)
    cdf = np.cumsum(np.array(rates) / total)
    return (cdf / cdf[-1]).tolist(), total


def make_ranges(lengths, wl):
    from mutation_simulator_b200._lib import MsRange
    cdf, total = type_cdf(wl["rates6"])
    arr = (MsRange * len(lengths))()
    for i, L in enumerate(lengths):
        a = arr[i]
        a.contig, a.start, a.stop, a.k, a.limit = i, 0, L - 1, int(((L - 1) + 1) * total), L   # mutator.py:225
        for t in range(7):
            a.cdf[t], a.minlen[t], a.maxlen[t] = cdf[t], wl["minlen"][t], wl["maxlen"][t]
    return arr


def lpt_partition(lengths, n):
    """Longest-processing-time bin packing of contigs onto n ranks."""
    from mutation_simulator_b200.distributed import lpt_partition as f
    return f(lengths, n)


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, polled through NVML every ~2 ms
    (the timed region is tens of milliseconds, too short for `nvidia-smi -lms`)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, device):
        self.device, self.samples, self.mask, self.max_mhz, self._stop, self._t = device, [], 0, None, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # torch's device order follows CUDA_VISIBLE_DEVICES; map through the UUID-less common case (same order)
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[device]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else device
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, str(e)

    def _loop(self):
        nv = self.nv
        while not self._stop:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {getattr(self, 'err', '')}"], "samples": 0}
        self._stop = True
        self._t.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for bit, n in self.REASONS.items() if self.mask & bit), "samples": len(self.samples)}


def reduce_over_ranks(values, op, world):
    """max / sum of a list of floats over the ranks (device-timed numbers are the max over ranks)."""
    if world == 1:
        return list(values)
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(values, dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return t.tolist()


def measured_peak():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------
# CPU arm: the oracle port
# ---------------------------------------------------------------------------------------
def cpu_run(lengths, names, wl, seed, threads):
    """One pass of the C oracle port (sample + walk + wrap + VCF) over a host genome."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import c_oracle
    c_oracle.lib()
    cdf, total = type_cdf(wl["rates6"])
    rng = np.random.default_rng(12345)
    seqs = [rng.choice(np.frombuffer(b"ACGT", np.uint8), L).tobytes() for L in lengths]

    def one(i):
        L = lengths[i]
        r = dict(start=0, stop=L - 1, k=int(L * total), cdf=cdf, minlen=wl["minlen"], maxlen=wl["maxlen"])
        fa, vcf, counts, _ = c_oracle.mutate_contig(seqs[i], names[i].encode(), names[i].encode(), 60, [r], [1] * 7, 1,
                                                    wl["titv"], seed + i)
        return len(fa) + len(vcf)

    def run():
        t0 = time.perf_counter()
        if threads > 1:
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(one, range(len(lengths))))
        else:
            for i in range(len(lengths)):
                one(i)
        return time.perf_counter() - t0
    return run


def scaled_sample(wl, target_bases):
    """Bounded sample of the workload: same contig count/shape, lengths scaled down."""
    lengths = wl["lengths"]
    tot = sum(lengths)
    if tot <= target_bases:
        return list(lengths), 1.0
    f = target_bases / tot
    return [max(1000, int(L * f)) for L in lengths], f


def reference_arm(args, wl, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    lengths, f = scaled_sample(wl, 400_000_000)
    names = wl["names"] or [f"ctg{i}" for i in range(len(lengths))]
    threads = min(threads, len(lengths))
    run = cpu_run(lengths, names, wl, 1, threads)
    for _ in range(max(1, min(args.warmup, 1))):
        run()
    ts = [run() for _ in range(args.steps)]
    t = float(np.mean(ts))
    val = sum(lengths) / t / 1e9
    line = {"impl": "reference", "metric": "mutated genome Gbp/s", "value": val, "unit": "Gbp/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": wl["desc"], "l2": "inputs larger than L2"},
            "cpu_baseline": {"value": val, "unit": "Gbp/s", "cores": threads, "kind": "port",
                             "sample": f"oracle/ms_oracle.c (C port of the reference's path; the reference itself is single-threaded "
                                       f"Python, 9.4e-4 Gbp/s in BASELINE.md) on the workload scaled x{f:.3f} = {sum(lengths)/1e6:.0f} Mbp, "
                                       f"one contig per thread"},
            "e2e": {"value": val, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def file_to_file(args, wl, eng, rank, world, lengths_all, names_all, total_bases):
    """The drop-in path end to end on files.  The input FASTA is the synthetic genome wrapped at 60 columns
    (built once, untimed, by applying an empty mutation table); the timed part is what a CLI user waits for."""
    import shutil
    import torch
    from mutation_simulator_b200 import Mutator, SimulationSettings, get_args, load_fasta
    from mutation_simulator_b200.engine import BUF_FASTA, Engine
    from mutation_simulator_b200.records import REC_DTYPE
    base = Path("/dev/shm") if Path("/dev/shm").is_dir() else Path(tempfile.gettempdir())
    d = base / f"ms_bench_{os.environ.get('MASTER_PORT', '0')}_{os.getppid() if world > 1 else os.getpid()}"
    src = d / "genome.fa"
    try:
        if rank == 0:
            d.mkdir(parents=True, exist_ok=True)
            g = Engine(int(os.environ.get("LOCAL_RANK", "0")))
            g.synth_genome(4242, lengths_all, [60] * len(lengths_all), [n.encode() for n in names_all],
                           [n.encode() for n in names_all], wl["n_fraction"], wl["telomere"])
            g.load_records(np.zeros(0, dtype=REC_DTYPE))
            g.apply()
            with open(src, "wb") as fh:
                fh.write(memoryview(g.download(BUF_FASTA)))
                fh.write(b"\n")
            g.close()
        if world > 1:
            torch.distributed.barrier()
        r6 = wl["rates6"]
        argv = [str(src), "-o", str(d / "out"), "-q", "--seed", "7", "args", "-sn", str(r6[0]), "-titv", str(wl["titv"])]
        for flag, rate, t in (("in", r6[1], 1), ("de", r6[2], 2), ("iv", r6[3], 3), ("du", r6[4], 4), ("tl", r6[5], 5)):
            if rate > 0:
                argv += [f"-{flag}", str(rate), f"-{flag}min", str(wl["minlen"][t]), f"-{flag}max", str(wl["maxlen"][t])]
        a = get_args(argv)
        a.device = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fasta = load_fasta(a.infile, device=a.device)   # FASTA ingest on the device (every rank; each keeps its contigs)
        t1 = time.perf_counter()
        sim = SimulationSettings.from_args(a, fasta, True)
        m = Mutator(a, fasta, sim)
        m.mutate()
        m.close()
        fasta.close()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        dt, parse = reduce_over_ranks([t2 - t0, t1 - t0], "max", world)
        out_bytes = (d / "out_ms.fa").stat().st_size + (d / "out_ms.vcf").stat().st_size if rank == 0 else 0
        return {"value": total_bases / dt / 1e9, "unit": "Gbp/s", "seconds": dt, "parse_seconds": parse,
                "in_bytes": src.stat().st_size if rank == 0 else 0, "out_bytes": out_bytes,
                "note": f"{base}: load_fasta + Mutator.mutate() writing .fa and .vcf, wall clock, max over ranks"}
    except Exception as e:  # noqa: BLE001 - the figure is auxiliary; never lose the main line over it
        return {"value": None, "error": f"{type(e).__name__}: {e}"}
    finally:
        if world > 1:
            try:
                torch.distributed.barrier()
            except Exception:  # noqa: BLE001
                pass
        if rank == 0:
            shutil.rmtree(d, ignore_errors=True)


# ---------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-f2f", action="store_true", help="skip the file-to-file measurement")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, wl, rank, world)
        return

    import torch
    import torch.distributed as dist
    from mutation_simulator_b200.engine import BUF_FASTA, BUF_VCF, Engine
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    lengths_all = wl["lengths"]
    names_all = wl["names"] or [f"ctg{i}" for i in range(len(lengths_all))]
    mine = lpt_partition(lengths_all, world)[rank] if world > 1 else list(range(len(lengths_all)))
    lengths = [lengths_all[i] for i in mine]
    names = [names_all[i].encode() for i in mine]
    my_bases = sum(lengths)
    total_bases = sum(lengths_all)

    eng = Engine(local_rank)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    eng.synth_genome(12345 + rank, lengths, [60] * len(lengths), names, names, wl["n_fraction"], wl["telomere"])
    ranges = make_ranges(lengths, wl)
    p_ti = wl["titv"] * (1 / (wl["titv"] + 1))
    eng.set_ranges_array(ranges, len(lengths), [1] * 7, 1, p_ti)

    def step(seed):
        eng.sample(seed)
        return eng.apply()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        step(1000 + w)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.stats()["kernel_launches"]
    stage_acc = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    t0 = time.perf_counter()
    sizes = (0, 0)
    for s in range(args.steps):
        sizes = step(2000 + s)
        if args.verbose or True:
            st = eng.stats()
            for k, v in st["stage_ms"].items():
                stage_acc[k] = stage_acc.get(k, 0.0) + v
    ev1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    st = eng.stats()
    launches = st["kernel_launches"] - launches0
    dev_ms = reduce_over_ranks([dev_ms], "max", world)[0]
    launches = int(reduce_over_ranks([float(launches)], "sum", world)[0])
    ms_per_step = dev_ms / args.steps
    value = total_bases / (ms_per_step * 1e-3) / 1e9

    # roofline of the dominant kernel (splice + SNP + line wrap -> FASTA image), rank 0's share
    recs_n = st["n_records"]
    fasta_bytes, vcf_bytes = sizes
    hdr_bytes = sum(len(n) + 2 for n in names)
    counts = st["counts"]
    splice_ms = stage_acc.get("splice_emit_fasta", 0.0) / args.steps
    tli_src = 25.5 * counts[6] if wl["maxlen"][5] == 50 else 0.0          # E[len] of the TLI gathers
    alg_bytes = my_bases + (fasta_bytes - hdr_bytes) + 32 * recs_n + tli_src
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (splice_ms * 1e-3) / 1e9 if splice_ms > 0 else 0.0
    traffic = None
    tf = REPO / "profiles" / "k_splice_traffic.json"
    if tf.exists() and args.workload == "c2" and world == 1:   # ncu --set full capture of this kernel on this workload
        t_ = json.loads(tf.read_text())
        traffic = t_["dram_bytes_read"] + t_["dram_bytes_write"]
    roofline = {"bound": "hbm", "kernel": "k_splice", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_launch": alg_bytes, "kernel_ms": splice_ms,
                "all_kernels_frac": ((my_bases + fasta_bytes + vcf_bytes + 64 * recs_n) / (ms_per_step * 1e-3) / 1e9) / peak}

    # e2e: HOST buffers through the C ABI
    e2e = None
    if not args.no_e2e:
        g = torch.empty(my_bases, dtype=torch.uint8, pin_memory=True)
        gn = g.numpy()
        gn[:] = eng.download_genome()
        fa_host = torch.empty(int(fasta_bytes * 1.02) + 4096, dtype=torch.uint8, pin_memory=True).numpy()
        vcf_host = torch.empty(int(vcf_bytes * 1.05) + 4096, dtype=torch.uint8, pin_memory=True).numpy()
        n_e2e = max(2, min(args.steps, 4))

        def e2e_serial(seed):     # the same work as four separate C-ABI calls, no overlap (reported for comparison)
            eng.upload_genome(gn, lengths, [60] * len(lengths), names, names, gid=mine)
            eng.set_ranges_array(ranges, len(lengths), [1] * 7, 1, p_ti)
            eng.sample(seed)
            fb, vb = eng.apply()
            eng.download(BUF_FASTA, fa_host)
            eng.download(BUF_VCF, vcf_host)
            return fb, vb

        def e2e_step(seed):       # ms_mutate_streamed: H2D / kernels / D2H of successive contig groups overlap
            eng.declare_genome(lengths, [60] * len(lengths), names, names, gid=mine)
            eng.set_ranges_array(ranges, len(lengths), [1] * 7, 1, p_ti)
            return eng.mutate_streamed(seed, gn, fa_host, vcf_host)
        e2e_serial(1)
        barrier()
        t0 = time.perf_counter()
        e2e_serial(2)
        barrier()
        serial_ms = reduce_over_ranks([(time.perf_counter() - t0) * 1e3], "max", world)[0]
        e2e_step(1)
        barrier()
        t0 = time.perf_counter()
        for s in range(n_e2e):
            fb, vb = e2e_step(3000 + s)
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        dt = reduce_over_ranks([dt], "max", world)[0]
        e2e = {"value": total_bases / dt / 1e9, "unit": "Gbp/s", "h2d_bytes_per_step": int(my_bases),
               "d2h_bytes_per_step": int(fb + vb), "steps": n_e2e, "ms_per_step": dt * 1e3,
               "serial_ms_per_step": serial_ms,
               "note": "per-rank bytes; pinned host genome in, FASTA image + VCF body out; one ms_mutate_streamed call per "
                       "step (copies of successive contig groups overlap the kernels); serial_ms_per_step = the same work "
                       "as upload/sample/apply/download calls without overlap"}

    # file-to-file: FASTA on (RAM-backed) disk -> load_fasta -> Mutator.mutate() -> *_ms.fa + *_ms.vcf on disk, wall clock
    f2f = None
    if not args.no_f2f:
        f2f = file_to_file(args, wl, eng, rank, world, lengths_all, names_all, total_bases)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        lens_s, f = scaled_sample(wl, 200_000_000)
        nm = wl["names"] or [f"ctg{i}" for i in range(len(lens_s))]
        run = cpu_run(lens_s, nm, wl, 1, 1)
        tcpu = run()
        cpu = {"value": sum(lens_s) / tcpu / 1e9, "unit": "Gbp/s", "cores": 1, "kind": "port",
               "sample": f"oracle/ms_oracle.c, 1 thread, workload scaled x{f:.3f} = {sum(lens_s)/1e6:.0f} Mbp in {tcpu:.1f} s "
                         f"(host has {os.cpu_count()} cores; the Python reference itself ran 9.4e-4 Gbp/s, BASELINE.md)"}

    if rank == 0:
        line = {"metric": "mutated genome Gbp/s", "value": value, "unit": "Gbp/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": wl["desc"], "partition": f"contigs LPT over {world} rank(s)",
                           "l2": "inputs (>=1 GB per rank) larger than the 126 MB L2; fresh seed every step",
                           "timing": "CUDA events on the engine's stream, max over ranks"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "file_to_file": f2f, "gpu_launches": int(launches), "clocks": clocks,
                "stage_ms": {k: v / args.steps for k, v in stage_acc.items()},
                "records_per_step": int(recs_n), "wall_ms_per_step": wall / args.steps * 1e3}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
