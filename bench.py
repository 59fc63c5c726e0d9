#!/usr/bin/env python
"""bench.py - throughput of the WGSassign hot path on B200 (contract in the task prompt).

Metric (BASELINE.json): site x individual x population evaluations per second, whole job.

    --config cfg3 (default; BASELINE.json configs[2], the single-GPU configuration the metric is quoted on):
        1,000,000 sites x 500 individuals x 10 populations per GPU; step = `--get_reference_af` + `--loo`.
    --config cfg4 (configs[3], the north-star target): 2,500,000 sites x 2,000 individuals x 20 populations per
        GPU - on 8 GPUs exactly the named 20 M-site job; step = `--get_pop_like` + `--get_reference_z_score` for
        all 2,000 individuals (class tally chained over the ranks in site order, keep mask, leave-one-out EM on the
        kept sites, moments).
    --config cfg5 (configs[4]): 5,000,000 x 1,000 x 8 on one GPU; step = `--ne_obs` (Fisher information,
        population + individual effective sample sizes) + `--get_pop_like` + `--get_em_mix`.
    --scaling weak (default): per-GPU sites fixed; strong: the configuration's site count divided over the ranks.

`value`  = M_total * N * K / step time with the GL (and depth) matrices resident in HBM.
`e2e`    = the same step through the host-buffer C ABI: pinned host matrices uploaded inside the timed region,
           results read back (cfg4 / cfg5: on a bounded site sample per GPU - 2 x 40 GB of host matrices per rank
           do not fit eight times in the box - which the block states).
`roofline` is the dominant kernel of the step against its BINDING bound; `kernels` carries every kernel family of
the step (algorithmic bytes / units from the library's own counters, CUDA-event time on the library's stream).
`parity` compares the GPU path with the oracle on the very slice `cpu_baseline` times.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config cfgX] [--scaling strong] [--sites M]
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "site*ind*pop evals/s (LOO+pop_like)"
UNIT = "evals/s"
MAF_ITER, MAF_TOLE = 200, 1e-4
SEED = 20261018
CONFIGS = {
    "cfg3": dict(n_ind=500, n_pop=10, sites=1_000_000, total_sites=1_000_000, with_ad=False,
                 name="cfg3: synthetic 1M sites x 500 individuals x 10 populations per GPU, --get_reference_af + --loo "
                      "(BASELINE.json configs[2])"),
    "cfg4": dict(n_ind=2000, n_pop=20, sites=2_500_000, total_sites=20_000_000, with_ad=True,
                 name="cfg4: synthetic 20M sites x 2,000 individuals x 20 populations on 8 GPUs (2.5M sites per GPU), "
                      "--get_pop_like + --get_reference_z_score (BASELINE.json configs[3])"),
    "cfg5": dict(n_ind=1000, n_pop=8, sites=5_000_000, total_sites=5_000_000, with_ad=False,
                 name="cfg5: synthetic 5M sites x 1,000 individuals x 8 populations, --ne_obs + --get_pop_like + --get_em_mix "
                      "(BASELINE.json configs[4])"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region.  The process is started BEFORE the warm-up
    (its start-up initialises the driver's management library, which stalls the CUDA calls of other processes for tens
    of milliseconds - inside a 0.5 s timed region that showed as 10..45 ms per step); only the samples stamped between
    begin() and end() are used."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_ms=200):
        self.gpu = gpu_index
        self.period_ms = period_ms
        self.proc = None
        self.path = None
        self.t_begin = self.t_end = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def begin(self):
        self.t_begin = time.time()

    def end(self):
        self.t_end = time.time()

    @staticmethod
    def _stamp(text):
        import datetime
        try:
            return datetime.datetime.strptime(text, "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        self.fh.close()
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 10:
                continue
            try:
                rows.append((self._stamp(f[0]), float(f[2]), float(f[3]), [nm for nm, v in zip(names, f[6:10]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.path)
        lo = (self.t_begin or 0.0) - 0.05
        hi = (self.t_end or float("inf")) + 0.05
        inside = [r for r in rows if r[0] is not None and lo <= r[0] <= hi]
        if not inside:                                             # a timed region shorter than the sampling period: nearest samples
            inside = [r for r in rows if r[0] is not None and lo - 0.2 <= r[0] <= hi + 0.2] or rows
        if inside:
            sm = [r[1] for r in inside]
            load = [x for x in sm if x >= 0.5 * max(sm)]
            reasons = sorted({nm for r in inside for nm in r[3]})
            out.update(sm_mhz=statistics.median(load), sm_max_mhz=max(r[2] for r in inside), reasons=reasons, samples=len(inside))
        return out


def pop_assignment(cfg):
    return ((np.arange(cfg["n_ind"]) * cfg["n_pop"]) // cfg["n_ind"]).astype(np.int32)


def ids_array(cfg):
    pop_of = pop_assignment(cfg)
    ids = np.empty((cfg["n_ind"], 2), dtype="U16")
    for i in range(cfg["n_ind"]):
        ids[i, 0] = "ind%d" % i
        ids[i, 1] = "pop%02d" % pop_of[i]
    return ids


# ------------------------------------------------------------------------------------------
# reference / CPU baseline leg (the ONLY place bench.py touches oracle/)
# ------------------------------------------------------------------------------------------
def _oracle():
    from oracle import oracle
    kind = "reference" if oracle.have_ref() else "port"
    return oracle, kind, oracle.kernels("ref" if kind == "reference" else "port")


def _kern_name(kind):
    return "the reference's compiled Cython" if kind == "reference" else "the oracle C port"


def cpu_step_cfg3(oracle, kern, L, IDs, threads):
    af, pops, its = oracle.reference_af(L, IDs, MAF_ITER, MAF_TOLE, threads, kern)
    a = af.copy()
    ll, _, lits = oracle.loo(L, a, IDs, threads, MAF_ITER, MAF_TOLE, kern=kern)
    return dict(af=af, its=its, ll=ll, lits=lits, af_after=a)


def cpu_cfg3(cfg, L, IDs, threads, target_s=12.0):
    """Time the reference's CPU path on a bounded slice; returns (cpu_baseline dict, sites used, oracle outputs)."""
    oracle, kind, kern = _oracle()
    m_probe = min(L.shape[0], 200)
    t0 = time.perf_counter()
    cpu_step_cfg3(oracle, kern, np.ascontiguousarray(L[:m_probe]), IDs, threads)
    dt = time.perf_counter() - t0
    m = int(min(L.shape[0], max(m_probe, m_probe * target_s / max(dt, 1e-3))))
    Ls = np.ascontiguousarray(L[:m])
    t0 = time.perf_counter()
    out = cpu_step_cfg3(oracle, kern, Ls, IDs, threads)
    dt = time.perf_counter() - t0
    evals = float(m) * cfg["n_ind"] * cfg["n_pop"]
    return ({"value": evals / dt, "unit": UNIT, "cores": threads, "kind": kind, "seconds": dt,
             "sample": "first %d sites x %d individuals x %d populations of the same synthetic matrix; --get_reference_af + --loo "
                       "through the restated drivers (oracle/oracle.py: the reference's loops, cost-equivalent) over %s kernels (-t %d)"
                       % (m, cfg["n_ind"], cfg["n_pop"], _kern_name(kind), threads)}, m, out)


def cpu_cfg4(cfg, L, AD, IDs, af, threads, n_pl=48, n_z=6):
    """pop_like on n_pl individuals and the reference z-score of n_z individuals over the given slice."""
    oracle, kind, kern = _oracle()
    m, K = L.shape[0], cfg["n_pop"]
    cols = np.repeat(np.arange(n_pl) * 2, 2) + np.tile([0, 1], n_pl)
    Lp = np.ascontiguousarray(L[:, cols])
    t0 = time.perf_counter()
    pl = oracle.assignLL(Lp, af, threads, kern)
    t_pl = time.perf_counter() - t0
    t0 = time.perf_counter()
    zr = oracle.zscore_reference(L, AD, IDs, MAF_ITER, MAF_TOLE, ind_start=0, ind_end=n_z, t=threads, kern=kern)
    t_z = time.perf_counter() - t0
    per_eval = t_pl / (float(m) * n_pl * K) + t_z / (float(m) * n_z * K)      # the step does both for every (site, individual)
    return ({"value": 1.0 / per_eval, "unit": UNIT, "cores": threads, "kind": kind, "seconds": t_pl + t_z,
             "pop_like_evals_per_s": float(m) * n_pl * K / t_pl, "zscore_site_ind_per_s": float(m) * n_z / t_z,
             "sample": "first %d sites of the same synthetic matrices: glassy.assignLL for %d individuals x %d populations (%.1f s) + the "
                       "--get_reference_z_score loop for %d individuals (%.1f s; pure-Python class tally / keep loops + %s kernels, -t %d); "
                       "value = evaluations per second of a step that does both for every (site, individual)"
                       % (m, n_pl, K, t_pl, n_z, t_z, _kern_name(kind), threads)}, dict(pl=pl, zr=zr, n_pl=n_pl, n_z=n_z))


def cpu_cfg5(cfg, L, IDs, af, threads):
    oracle, kind, kern = _oracle()
    m, n, K = L.shape[0], cfg["n_ind"], cfg["n_pop"]
    t0 = time.perf_counter()
    f_obs, ne_obs = oracle.fisher_obs(L, af, IDs, threads, kern)
    ne_ind = oracle.fisher_obs_ind(L, af, IDs, threads, kern)
    t_f = time.perf_counter() - t0
    n_pl = 48
    cols = np.repeat(np.arange(n_pl) * 2, 2) + np.tile([0, 1], n_pl)
    t0 = time.perf_counter()
    pl = oracle.assignLL(np.ascontiguousarray(L[:, cols]), af, threads, kern)
    t_pl = time.perf_counter() - t0
    per_eval = t_f / (float(m) * n * K) + t_pl / (float(m) * n_pl * K)
    return ({"value": 1.0 / per_eval, "unit": UNIT, "cores": threads, "kind": kind, "seconds": t_f + t_pl,
             "fisher_site_ind_per_s": float(m) * n / t_f,
             "sample": "first %d sites: fisher.fisher_obs + fisher_obs_ind for all %d individuals (%.1f s) + glassy.assignLL for %d "
                       "individuals (%.1f s), %s kernels, -t %d" % (m, n, t_f, n_pl, t_pl, _kern_name(kind), threads)},
            dict(f_obs=f_obs, ne_obs=ne_obs, ne_ind=ne_ind, pl=pl, n_pl=n_pl))


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the box's host cores, each step a
    bounded sample of the arm's workload.  Under torchrun rank 0 alone works."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from wgsassign_b200 import synth
    cfg = CONFIGS[args.config]
    oracle, kind, kern = _oracle()
    threads = os.cpu_count() or 1
    IDs = ids_array(cfg)
    N, K = cfg["n_ind"], cfg["n_pop"]
    if args.config == "cfg3":
        probe = synth.synth(200, N, K, seed=SEED, with_ad=False)["L"]
        t0 = time.perf_counter()
        cpu_step_cfg3(oracle, kern, probe, IDs, threads)
        dt = time.perf_counter() - t0
        m = int(max(200, min(20000, 200 * 8.0 / max(dt, 1e-3))))      # ~8 s of CPU work per step
        L = synth.synth(m, N, K, seed=SEED, with_ad=False)["L"]
        step = lambda: cpu_step_cfg3(oracle, kern, L, IDs, threads)
        evals = float(m) * N * K
        what = "--get_reference_af + --loo per step through the restated drivers (the reference's loops, cost-equivalent)"
        value_of = lambda dt: evals / dt
    else:
        m = 3000 if args.config == "cfg4" else 20000
        d = synth.synth(m, N, K, seed=SEED, with_ad=cfg["with_ad"])
        L = d["L"]
        af, _, _ = oracle.reference_af(L, IDs, MAF_ITER, MAF_TOLE, threads, kern)
        box = {}
        if args.config == "cfg4":
            step = lambda: box.update(cb=cpu_cfg4(cfg, L, d["AD"], IDs, af, threads, n_pl=16, n_z=2)[0])
            what = "--get_pop_like + --get_reference_z_score on an individual subset per step"
        else:
            step = lambda: box.update(cb=cpu_cfg5(cfg, L, IDs, af, threads)[0])
            what = "--ne_obs + --get_pop_like per step"
        value_of = lambda dt: box["cb"]["value"]
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    vals = []
    for _ in range(args.steps):
        t1 = time.perf_counter()
        step()
        vals.append(value_of(time.perf_counter() - t1))
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    val = float(np.mean(vals)) if args.config != "cfg3" else float(m) * N * K / dt
    sample = ("%d sites x %d individuals x %d populations (NumPy generator, same model as the GPU arm); %s over %s kernels, -t %d"
              % (m, N, K, what, _kern_name(kind), threads))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"] + " (bounded site slice on the CPU)", "sample_sites": m, "individuals": N, "populations": K,
                       "maf_iter": MAF_ITER, "maf_tole": MAF_TOLE},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------
# GPU leg
# ------------------------------------------------------------------------------------------
def ncu_traffic(name, sites):
    """DRAM bytes per launch of a kernel family from the committed ncu captures, scaled to `sites`."""
    for fn in ("ncu_traffic_r2.json", "ncu_traffic_r1.json"):
        try:
            k = json.load(open(os.path.join(ROOT, "profiles", fn)))["kernels"][name]
            return k["dram_bytes_per_launch"] * sites / k["sites"]
        except Exception:
            continue
    return None


def family(ctx, name, peak_gbs, sites=None, with_traffic=True):
    t = ctx.timing_get(name)
    if t["launches"] == 0 or t["ms"] <= 0:
        return None
    sec = t["ms"] * 1e-3
    out = {"launches": t["launches"], "ms_total": t["ms"], "ms_per_launch": t["ms"] / t["launches"],
           "algorithmic_gb": t["bytes"] / 1e9, "achieved_gbs": t["bytes"] / 1e9 / sec, "hbm_frac": t["bytes"] / 1e9 / sec / peak_gbs,
           "units": t["units"], "units_per_s": t["units"] / sec}
    if sites and with_traffic:
        out["algorithmic_bytes_per_launch"] = t["bytes"] / t["launches"]
        out["ncu_dram_bytes_per_launch"] = ncu_traffic(name, sites)
    return out


def ingest_probe(cfg, ctx_device):
    """Throughput of the text ingestion (timed separately from the metric, as BASELINE.json asks): a synthetic Beagle file
    and allele-depth file of this configuration's width - ~200 MB of text each, gzip level 1, built from one 200-row
    block repeated as separate gzip members (rows are longer than deflate's window, so the ratio is that of real data) -
    through the streaming readers, and the Beagle file once more straight into a resident context (parse + upload
    overlapped, what the CLI does)."""
    import gzip, io, tempfile
    from wgsassign_b200 import reader, session, synth
    n = cfg["n_ind"]
    rows_blk = 200
    d = synth.synth(rows_blk, n, cfg["n_pop"], seed=7, with_ad=True)
    g2 = np.abs(np.round(1.0 - d["L"][:, 0::2].astype(np.float64) - d["L"][:, 1::2], 6))
    cells = np.empty((rows_blk, 3 * n))
    cells[:, 0::3], cells[:, 1::3], cells[:, 2::3] = d["L"][:, 0::2], d["L"][:, 1::2], g2
    buf = io.StringIO()
    np.savetxt(buf, cells, fmt="%.6f", delimiter="\t")
    body = "".join("chr1_%d\t0\t1\t%s\n" % (i, ln) for i, ln in enumerate(buf.getvalue().splitlines())).encode()
    reps = max(1, int(200e6 / len(body)))
    header = ("marker\tallele1\tallele2" + "".join("\t%s\t%s\t%s" % (("s%d" % i,) * 3) for i in range(n)) + "\n").encode()
    tmp = tempfile.mkdtemp()
    pb, pa = os.path.join(tmp, "x.beagle.gz"), os.path.join(tmp, "x.ad.txt.gz")
    with open(pb, "wb") as fh:
        fh.write(gzip.compress(header, 1))
        member = gzip.compress(body, 1)
        for _ in range(reps):
            fh.write(member)
    abuf = io.StringIO()
    np.savetxt(abuf, d["AD"], fmt="%d")
    abody = abuf.getvalue().encode()
    areps = max(1, int(100e6 / len(abody)))
    with open(pa, "wb") as fh:
        member = gzip.compress(abody, 1)
        for _ in range(areps):
            fh.write(member)
    threads = os.cpu_count() or 1
    out = {"threads": threads, "note": "timed separately from the metric; parse = (float)atof-exact conversion into pinned memory"}
    t0 = time.perf_counter()
    L, _, sites = reader.readBeagle(pb, threads)
    out["beagle"] = dict(reader.last_stats["beagle"], shape=[int(L.shape[0]), int(L.shape[1])])
    del L
    session.reset()
    t0 = time.perf_counter()
    ctx, L, _, _ = session.stream_context(pb, pop_assignment(cfg), cfg["n_pop"], threads)
    ctx.upload_wait()
    wall = time.perf_counter() - t0
    st = reader.last_stats["beagle"]
    out["beagle_to_device"] = {"wall_s": wall, "compressed_mb_per_s": st["compressed_bytes"] / 1e6 / wall,
                               "uncompressed_mb_per_s": st["uncompressed_bytes"] / 1e6 / wall, "matrix_gb_per_s": L.nbytes / 1e9 / wall,
                               "path": "readBeagle(on_block) + wgs_upload_gl_begin/rows/end: the upload of block b overlaps the parsing of block b+1"}
    session.reset()
    # the same text as BGZF (what ANGSD writes): the members are inflated in parallel
    pz = os.path.join(tmp, "x.bgzf.beagle.gz")
    with open(pz, "wb") as fh:
        member = synth.bgzf_compress(body)[:-28]                 # without the end-of-file member
        fh.write(synth.bgzf_compress(header)[:-28])
        for _ in range(reps):
            fh.write(member)
        fh.write(synth.bgzf_compress(b"")[-28:])
    L, _, _ = reader.readBeagle(pz, threads)
    out["beagle_bgzf"] = dict(reader.last_stats["beagle"], shape=[int(L.shape[0]), int(L.shape[1])])
    del L
    os.unlink(pz)
    AD = reader.readAD(pa, threads)
    out["allele_depths"] = dict(reader.last_stats["ad"], shape=[int(AD.shape[0]), int(AD.shape[1])], dtype=str(AD.dtype))
    for f in (pb, pa):
        os.unlink(f)
    os.rmdir(tmp)
    return out


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--sites", type=int, default=None, help="sites per GPU (weak) or in total (strong); default: the configuration's")
    ap.add_argument("--clock-period-ms", type=int, default=200, help="nvidia-smi sampling period during the run (0: no sampling; diagnosis only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the per-kernel extras")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 10 if args.config == "cfg3" else 3
    if args.warmup < 3 and args.impl == "b200" and not os.environ.get("WGS_BENCH_ALLOW_SHORT"):
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    cfg = CONFIGS[args.config]
    N_IND, N_POP = cfg["n_ind"], cfg["n_pop"]

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    from wgsassign_b200 import _lib, dist, mixture
    numa = dist.bind_near_gpu(local) if world > 1 else "numa: single process, not bound"   # before any pinned allocation
    if world > 1:
        import torch.distributed as td
        # keep rank 0's stdout to the one JSON line: NCCL prints its version banner to stdout at
        # communicator creation for any NCCL_DEBUG level >= VERSION, so create it with fd 1 -> stderr
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            td.init_process_group("nccl", device_id=torch.device("cuda", local))
            td.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    if args.scaling == "strong":
        M_total = args.sites or cfg["total_sites"]
        lo, hi = dist.shard_range(M_total, rank, world)
        M_local, offset = hi - lo, lo
    else:
        M_local = args.sites or cfg["sites"]
        M_total, offset = M_local * world, rank * M_local
    if world > 1:
        dist.enable(M_total, offset, device=torch.device("cuda", local))

    hbm_peak, peak_src, sm_max = peaks()
    ctx = _lib.Context(local)
    pop_of = pop_assignment(cfg)
    ctx.set_pops(pop_of, N_POP)
    dist.attach(ctx)           # must precede synth: the generator is keyed by global site index
    ctx.synth(M_local, N_IND, seed=SEED, with_ad=cfg["with_ad"])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as td
            td.barrier()
        torch.cuda.synchronize()

    state = {}
    if args.config != "cfg3":
        state["af"], state["af_its"] = ctx.ref_af(MAF_ITER, MAF_TOLE)            # the AF file these modes read (untimed)
        state["mix_ids"] = np.stack([np.arange(N_IND).astype("U8"), np.array(["h%d" % (i % 3) for i in range(N_IND)])], axis=1)

    def step():
        """One pass of the configuration's CLI modes over the resident matrices."""
        if args.config == "cfg3":
            # the AF matrix stays on the device between the two operators
            af, its = ctx.ref_af(MAF_ITER, MAF_TOLE, download=False)
            ll, _, lits = ctx.loo_partial(af, MAF_ITER, MAF_TOLE)
            dist.combine(ctx, ll)
            return dict(af=af, ll=ll, its=its, lits=lits)
        if args.config == "cfg4":
            pl = ctx.pop_like_partial(state["af"])
            dist.combine(ctx, pl)
            zr = ctx.zscore(1, None, 0, False, 0, N_IND, MAF_ITER, MAF_TOLE)
            return dict(ll=pl, zr=zr)
        f_obs, ne_obs, ind = ctx.fisher_partial(state["af"])
        dist.combine(ctx, ind)
        pl = ctx.pop_like_partial(state["af"])
        dist.combine(ctx, pl)
        mix = mixture.em_mix(pl.astype(np.float32), state["mix_ids"], 200) if rank == 0 else None
        return dict(ll=pl, ne_ind=ind / float(M_total), f_obs=f_obs, ne_obs=ne_obs, mix=mix)

    sampler = ClockSampler(local, args.clock_period_ms)
    if rank == 0 and args.clock_period_ms > 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    sampler.begin()
    ctx.timing_reset(True)
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = step()
    barrier()
    dt = time.perf_counter() - t0
    sampler.end()
    launches = ctx.launch_count() - l0
    fam_names = {"cfg3": ("loo_em", "loo_first", "em_pop", "loo_like", "loo_like_aux", "loo_pack", "em_resolve"),
                 "cfg4": ("pop_like", "pop_like_aux", "ztally", "zkeep", "zmoments", "loo_em", "loo_first", "loo_pack", "em_resolve"),
                 "cfg5": ("fisher", "pop_like", "pop_like_aux")}[args.config]
    fam = {k: family(ctx, k, hbm_peak, M_local) for k in fam_names}
    gaps = {k: ctx.timing_get(k) for k in ("gap", "gap_long")}     # idle time of the stream between consecutive launches
    ctx.timing_reset(False)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        import torch.distributed as td
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        td.all_reduce(tt, op=td.ReduceOp.MAX)
        dt = float(tt.item())
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        td.all_reduce(lt, op=td.ReduceOp.SUM)
        launches = int(lt.item())
    ms_step = dt / args.steps * 1e3
    evals_step = float(M_total) * N_IND * N_POP
    value = evals_step / (dt / args.steps)
    assign_ok = float(np.mean(np.argmax(res["ll"], 1) == pop_of))

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        M_e = M_local if args.config == "cfg3" else min(M_local, 250_000 if args.config == "cfg4" else 500_000)
        Lh = _lib.pinned_empty((M_e, 2 * N_IND), np.float32)
        ADh = _lib.pinned_empty((M_e, 2 * N_IND), np.int32) if cfg["with_ad"] else None
        chunk = 50_000
        for s0 in range(0, M_e, chunk):
            n = min(chunk, M_e - s0)
            if ADh is None:
                Lh[s0:s0 + n] = ctx.download(s0, n)
            else:
                Lh[s0:s0 + n], ADh[s0:s0 + n] = ctx.download(s0, n, want_ad=True)
        af_e = None if args.config == "cfg3" else np.ascontiguousarray(state["af"][:M_e])
        Me_total, off_e = M_e * world, rank * M_e
        full_res = res
        ctx.close()                                   # the resident matrices go: the e2e arm starts from host memory
        if world > 1:
            dist.enable(Me_total, off_e, device=torch.device("cuda", local))
        ctx2 = _lib.Context(local)
        af_pin = _lib.pinned_empty((M_e, N_POP), np.float32)      # host buffer the allele frequencies come back into

        def e2e_step():
            """What the CLI does on freshly parsed matrices.  cfg3: queue the upload (one strided DMA per population
            slab), then ONE fused call whose leave-one-out EM starts on the first slab while the later ones are still
            crossing PCIe; AF (for <out>.pop_af.npy) and likelihoods come back.  cfg4 / cfg5: upload, then the modes."""
            ctx2.set_pops(pop_of, N_POP)
            dist.attach(ctx2)
            if args.config == "cfg3":
                ctx2.upload_gl_async(Lh)
                af_, its_, ll_, _, lits_, _ = ctx2.ref_af_loo(MAF_ITER, MAF_TOLE, af_out=af_pin)
                dist.combine(ctx2, ll_)
                return dict(af=af_, ll=ll_)
            ctx2.upload_gl(Lh)
            if args.config == "cfg4":
                ctx2.upload_ad(ADh)
                pl = ctx2.pop_like_partial(af_e)
                dist.combine(ctx2, pl)
                zr = ctx2.zscore(1, None, 0, False, 0, N_IND, MAF_ITER, MAF_TOLE)
                return dict(ll=pl, zr=zr)
            f_obs, ne_obs, ind = ctx2.fisher_partial(af_e)
            dist.combine(ctx2, ind)
            pl = ctx2.pop_like_partial(af_e)
            dist.combine(ctx2, pl)
            return dict(ll=pl, f_obs=f_obs)
        r2 = e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            r2 = e2e_step()
        barrier()
        dte = time.perf_counter() - t0
        if world > 1:
            import torch.distributed as td
            tt = torch.tensor([dte], dtype=torch.float64, device="cuda")
            td.all_reduce(tt, op=td.ReduceOp.MAX)
            dte = float(tt.item())
        h2d = Lh.nbytes + pop_of.nbytes + (ADh.nbytes if ADh is not None else 0) + (af_e.nbytes * (2 if args.config == "cfg5" else 1) if af_e is not None else 0)
        if args.config == "cfg3":
            d2h = r2["af"].nbytes + 2 * r2["ll"].nbytes + 4 * (N_IND + N_POP)
        elif args.config == "cfg4":
            d2h = r2["ll"].nbytes + 32 * N_IND
        else:
            d2h = 2 * r2["f_obs"].nbytes + r2["ll"].nbytes + 8 * N_IND
        e2e = {"value": float(Me_total) * N_IND * N_POP / (dte / n_e2e), "unit": UNIT, "h2d_bytes_per_step": int(h2d * world),
               "d2h_bytes_per_step": int(d2h * world), "ms_per_step": dte / n_e2e * 1e3, "steps": n_e2e,
               "sites_per_gpu": M_e,
               "h2d_gbs_per_gpu": h2d / 1e9 / (dte / n_e2e),
               "path": {"cfg3": "wgs_upload_gl_async + wgs_ref_af_loo (leave-one-out EM of population k overlaps the upload of k+1..)",
                        "cfg4": "wgs_upload_gl + wgs_upload_ad + wgs_pop_like_partial + wgs_zscore(mode 1)",
                        "cfg5": "wgs_upload_gl + wgs_fisher_partial + wgs_pop_like_partial"}[args.config]}
        if args.config == "cfg3":
            e2e["identical_to_resident"] = bool(np.array_equal(r2["ll"], full_res["ll"]))
            # the upload alone, every rank at once: one strided DMA per population slab (what e2e uses) against the chunked
            # contiguous copy + permutation kernel - shows what the host memory system gives N concurrent uploads
            up = {}
            for name, fn in (("slab_dma", lambda: (ctx2.upload_gl_async(Lh), ctx2.upload_wait())), ("chunked_repack", lambda: ctx2.upload_gl(Lh))):
                ctx2.set_pops(pop_of, N_POP)
                fn()
                barrier()
                t0 = time.perf_counter()
                fn()
                barrier()
                dtu = time.perf_counter() - t0
                if world > 1:
                    tt = torch.tensor([dtu], dtype=torch.float64, device="cuda")
                    td.all_reduce(tt, op=td.ReduceOp.MAX)
                    dtu = float(tt.item())
                up[name] = {"ms": dtu * 1e3, "gb_per_s_per_gpu": Lh.nbytes / 1e9 / dtu, "gb_per_s_all": Lh.nbytes * world / 1e9 / dtu}
            e2e["upload_only"] = up
            dist.attach(ctx2)
            ctx2.upload_gl(Lh)
        else:
            e2e["note"] = ("bounded sample: %d of the %d sites per GPU (the full host matrices, %.0f GB per rank, do not fit N times in host memory)"
                           % (M_e, M_local, (Lh.nbytes + (ADh.nbytes if ADh is not None else 0)) * M_local / M_e / 1e9))
        ctx = ctx2
        M_res = M_e
    else:
        M_res = M_local

    # ---- extras: the other kernels of the path on the resident matrix (untimed for `value`) ----
    extra = {}
    if not args.no_extra:
        af = state.get("af")
        if af is None or af.shape[0] != M_res:
            af, _ = ctx.ref_af(MAF_ITER, MAF_TOLE)
        ctx.pop_like_partial(af); ctx.fisher_partial(af)        # untimed first use (module load, pool growth)
        ctx.timing_reset(True)
        for _ in range(3):
            ctx.pop_like_partial(af)
            ctx.fisher_partial(af)
        extra = {k: family(ctx, k, hbm_peak, M_res) for k in ("pop_like", "fisher")}
        aux = ctx.timing_get("pop_like_aux")                    # coefficient rows + population-only log term of the ratio form
        if extra.get("pop_like") and aux["launches"]:
            pl_ = extra["pop_like"]
            pl_["aux_ms_per_call"] = aux["ms"] / pl_["launches"]
            pl_["ms_per_call_with_aux"] = (pl_["ms_total"] + aux["ms"]) / pl_["launches"]
            pl_["hbm_frac_with_aux"] = pl_["algorithmic_gb"] / ((pl_["ms_total"] + aux["ms"]) * 1e-3) / hbm_peak
        if args.config == "cfg4" and M_res <= 600_000:
            # the --loo operator at this shape (not part of the step): loo_em / loo_like / loo_pack figures
            ctx.timing_reset(True)
            ctx.loo_partial(af.copy(), MAF_ITER, MAF_TOLE)
            extra.update({"loo_op_" + k: family(ctx, k, hbm_peak, M_res, False) for k in ("loo_em", "loo_like", "loo_like_aux", "loo_pack", "loo_first")})
        for mode, nm in ((0, "stream_flat"), (1, "stream_slab")):       # what a pure read of the same matrix achieves
            ms, nb = ctx.debug_stream(mode)
            extra[nm] = {"ms": ms, "gb": nb / 1e9, "achieved_gbs": nb / 1e9 / (ms * 1e-3)}
        ctx.timing_reset(False)
        if world == 1:
            try:
                extra["ingest"] = ingest_probe(cfg, local)
            except Exception as e:                              # never let the side measurement take the line down
                extra["ingest"] = {"error": repr(e)}

    # ---- CPU baseline + parity on the same slice (rank 0, one GPU) ----
    cpu_b, parity = None, None
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        IDs = ids_array(cfg)
        pctx = _lib.Context(local)
        pctx.set_pops(pop_of, N_POP)
        if args.config == "cfg3":
            Ls = ctx.download(0, min(M_res, 20000))
            cpu_b, m_used, o = cpu_cfg3(cfg, Ls, IDs, threads)
            Ls = np.ascontiguousarray(Ls[:m_used])
            pctx.upload_gl(Ls)
            af_g, its_g = pctx.ref_af(MAF_ITER, MAF_TOLE)
            a_g = af_g.copy()
            ll_g, _, lits_g = pctx.loo_partial(a_g, MAF_ITER, MAF_TOLE)
            ok = np.isfinite(o["ll"])
            parity = {"against": "oracle (%s kernels) on the cpu_baseline slice: %d sites x %d x %d" % (cpu_b["kind"], m_used, N_IND, N_POP),
                      "af_max_abs": float(np.max(np.abs(af_g - o["af"]))), "af_after_loo_max_abs": float(np.max(np.abs(a_g - o["af_after"]))),
                      "ll_max_rel": rel_err(ll_g[ok], o["ll"][ok]),
                      "argmax_equal": bool(np.array_equal(np.argmax(ll_g, 1), np.argmax(o["ll"], 1))),
                      "iters_equal": bool(list(its_g) == list(o["its"]) and list(lits_g) == list(o["lits"])),
                      "tolerances": {"af": 1e-5, "ll_rel": 1e-6}}
        elif args.config == "cfg4":
            m_s = min(M_res, 20000)
            Ls, ADs = ctx.download(0, m_s, want_ad=True)
            pctx.upload_gl(Ls); pctx.upload_ad(ADs)
            af_s, _ = pctx.ref_af(MAF_ITER, MAF_TOLE)
            cpu_b, o = cpu_cfg4(cfg, Ls, ADs, IDs, af_s, threads)
            pl_g = pctx.pop_like_partial(af_s)[:o["n_pl"]]
            zr_g = pctx.zscore(1, None, 0, False, 0, o["n_z"], MAF_ITER, MAF_TOLE)
            comp = 0.0
            for g_, r_ in zip(zr_g, o["zr"]):
                for a_, b_ in ((g_.w_obs, r_["w_obs"]), (g_.z_mu, r_["z_mu"]), (g_.z_var, r_["z_var"])):
                    comp = max(comp, abs(float(a_) - float(b_)) / abs(float(b_)))
            parity = {"against": "oracle (%s kernels) on the cpu_baseline slice: %d sites; pop_like of %d individuals, reference z-score of %d"
                                 % (cpu_b["kind"], m_s, o["n_pl"], o["n_z"]),
                      "ll_max_rel": rel_err(pl_g, o["pl"]), "argmax_equal": bool(np.array_equal(np.argmax(pl_g, 1), np.argmax(o["pl"], 1))),
                      "z_loci_kept_equal": bool([g_.loci_kept for g_ in zr_g] == [r_["loci_kept"] for r_ in o["zr"]]),
                      "z_em_iters_equal": bool([g_.em_iters for g_ in zr_g] == [r_["em_iter"] for r_ in o["zr"]]),
                      "z_components_max_rel": comp, "z_max_abs": float(max(abs(float(g_.z) - float(r_["z"])) for g_, r_ in zip(zr_g, o["zr"]))),
                      "tolerances": {"ll_rel": 1e-6, "z_components_rel": 1e-6}}
        else:
            m_s = min(M_res, 12000)
            Ls = ctx.download(0, m_s)
            pctx.upload_gl(Ls)
            af_s, _ = pctx.ref_af(MAF_ITER, MAF_TOLE)
            cpu_b, o = cpu_cfg5(cfg, Ls, IDs, af_s, threads)
            f_g, ne_g, ind_g = pctx.fisher_partial(af_s)
            pl_g = pctx.pop_like_partial(af_s)[:o["n_pl"]]
            scale = np.abs(o["f_obs"]).max(0)
            parity = {"against": "oracle (%s kernels) on the cpu_baseline slice: %d sites x %d x %d" % (cpu_b["kind"], m_s, N_IND, N_POP),
                      "fisher_obs_max_err_over_column_scale": float(np.max(np.abs(f_g - o["f_obs"]) / scale)),
                      "ne_ind_max_rel": rel_err(ind_g / m_s, o["ne_ind"]), "ll_max_rel": rel_err(pl_g, o["pl"]),
                      "argmax_equal": bool(np.array_equal(np.argmax(pl_g, 1), np.argmax(o["pl"], 1)))}
        pctx.close()

    if world > 1:
        import torch.distributed as td
        td.barrier()
        td.destroy_process_group()
    if rank != 0:
        return 0

    sm_mhz = (clocks or {}).get("sm_mhz") or sm_max
    # Issue-rate ceilings MEASURED on this pool's B200 (scripts/microbench/, profiles/issue_rates_r1.txt,
    # profiles/mufu_rate_r1.txt), scaled by the clock this run sustained:
    #   packed FP32 (FFMA2): 1.687e13 lane-instructions/s at 1,965 MHz = 58.0 per clock per SM;
    #   MUFU.RCP: 15.48 per clock per SM.
    # loo_em_step5 spends, per 4 posterior evaluations, 7 FFMA2 + 2 FFMA (= 8 packed-issue equivalents on the FP32 pipe)
    # and 2 MUFU.RCP: 2.0 packed lane-instructions and 0.5 reciprocals per evaluation.  Both pipes are loaded to within
    # 5 % of each other; the FP32 pipe is the (slightly) tighter one and is the binding bound.
    fp32_peak = 148 * 58.0 * sm_mhz * 1e6 / 2.0
    mufu_peak = 148 * 15.48 * sm_mhz * 1e6 / 0.5
    le = fam.get("loo_em")
    roofline = None
    if le:
        roofline = {"kernel": "loo_em_step5_kernel", "bound": "fp32_issue", "achieved": le["units_per_s"], "peak": fp32_peak,
                    "unit": "posterior evals/s", "frac": le["units_per_s"] / fp32_peak,
                    "traffic": le.get("ncu_dram_bytes_per_launch"),
                    "peak_source": "measured packed-FP32 issue rate, 58.0 lane-instructions/clk/SM at the sampled clock (profiles/issue_rates_r1.txt) "
                                   "/ 2.0 per evaluation",
                    "instructions_per_eval": {"ffma2": 1.75, "ffma": 0.5, "mufu_rcp": 0.5, "packed_fp32_issue_equivalents": 2.0},
                    "sm_mhz": sm_mhz, "mufu_rcp_frac": le["units_per_s"] / mufu_peak,
                    "ms_per_launch": le["ms_per_launch"], "launches": le["launches"], "share_of_step": le["ms_total"] / (dt * 1e3),
                    "hbm": {"bound": "hbm (not binding: each packed row group is read once per iteration and re-used for n^2 evaluations)",
                            "achieved": le["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": le["hbm_frac"], "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": le.get("algorithmic_bytes_per_launch")},
                    "inner_loop_ceiling": "6.3e12 evals/s for the bare inner loop from shared memory (profiles/loo_quad_rate_r1.txt)"}
    elif fam.get("fisher"):
        fk = fam["fisher"]
        roofline = {"kernel": "fisher2_kernel", "bound": "hbm", "achieved": fk["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": fk["hbm_frac"], "traffic": fk.get("ncu_dram_bytes_per_launch"), "peak_source": peak_src,
                    "ms_per_launch": fk["ms_per_launch"], "launches": fk["launches"], "share_of_step": fk["ms_total"] / (dt * 1e3)}
    conf = {"workload": cfg["name"], "sites_per_gpu": M_local, "sites_total": M_total, "individuals": N_IND, "populations": N_POP,
            "maf_iter": MAF_ITER, "maf_tole": MAF_TOLE, "parallelism": "site-sharded x%d" % world,
            "l2": "inputs (%.1f GB GL per GPU) exceed the 126 MB L2, no flush needed" % (M_local * N_IND * 8 / 1e9),
            "timer": "host clock around blocking C-ABI calls, device-synchronised + barrier on both sides, max over ranks; "
                     "kernels timed with CUDA events on the library's stream",
            "self_assignment_rate": assign_ok, "host_binding_rank0": numa}
    if args.config == "cfg3":
        conf.update(em_iters_ref=[int(x) for x in res["its"]], em_iters_loo_minmax=[int(np.min(res["lits"])), int(np.max(res["lits"]))])
    if args.config == "cfg4":
        zz = np.array([r.z for r in res["zr"]], np.float64)
        conf.update(z_mean=float(np.nanmean(zz)), z_sd=float(np.nanstd(zz)), z_loci_kept_mean_frac=float(np.mean([r.loci_kept for r in res["zr"]]) / M_total),
                    z_em_iters_minmax=[int(min(r.em_iters for r in res["zr"])), int(max(r.em_iters for r in res["zr"]))])
    if args.config == "cfg5":
        conf.update(ne_ind_mean=float(np.mean(res["ne_ind"])), em_mix_rows=None if res["mix"] is None else [list(map(str, r)) for r in res["mix"]])
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": conf, "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": roofline,
            "kernels": {**fam, **extra}, "cpu_baseline": cpu_b, "parity": parity,
            "stream_idle": {"ms_per_step": gaps["gap"]["ms"] / args.steps, "gaps_per_step": gaps["gap"]["launches"] / args.steps,
                            "long_ms_per_step": gaps["gap_long"]["ms"] / args.steps, "long_gaps_per_step": gaps["gap_long"]["launches"] / args.steps,
                            "note": "time the stream sat idle between consecutive kernels of the timed steps (long: gaps above 50 us)"}}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
