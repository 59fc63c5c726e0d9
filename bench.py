#!/usr/bin/env python
"""bench.py - throughput of the WGSassign hot path on B200 (contract in the task prompt).

Metric (BASELINE.json): site x individual x population evaluations per second for
LOO + pop_like.  One "step" = one pass of `--get_reference_af` + `--loo` (per-population EM,
leave-one-out EM for every individual, leave-one-out likelihoods with the reference's column
overwrite order) over one batch of synthetic Beagle-shaped input:

    workload cfg3 = BASELINE.json configs[2]: 1,000,000 sites x 500 individuals x 10 populations
    per GPU (weak scaling: N GPUs hold N x 1M sites, sharded by site; the EM stop rule is
    global, so ranks exchange one float64 per EM problem per iteration).

`value`  = M_total * N * K / step time with the GL matrix resident in HBM.
`e2e`    = the same step through the host-buffer C ABI: the pinned host matrix is uploaded
           (H2D, repack) inside the timed region and the results are read back.
`roofline` is for the dominant kernel (loo_em_step); `kernels` carries the same figures for
every other kernel of the path, `cpu_baseline` the reference's own compiled kernels
(oracle/_ref) or the oracle port on the host cores over a bounded slice of the same data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--sites M]
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "site*ind*pop evals/s (LOO+pop_like)"
UNIT = "evals/s"
N_IND, N_POP = 500, 10
SITES_PER_GPU = 1_000_000
MAF_ITER, MAF_TOLE = 200, 1e-4
SEED = 20261018


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            load = [x for x in sm if x >= 0.5 * max(sm)]
            out.update(sm_mhz=statistics.median(load), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def pop_assignment():
    return ((np.arange(N_IND) * N_POP) // N_IND).astype(np.int32)


def ids_array():
    pop_of = pop_assignment()
    ids = np.empty((N_IND, 2), dtype="U16")
    for i in range(N_IND):
        ids[i, 0] = "ind%d" % i
        ids[i, 1] = "pop%02d" % pop_of[i]
    return ids


# ------------------------------------------------------------------------------------------
# reference / CPU baseline leg
# ------------------------------------------------------------------------------------------
def cpu_step(oracle, kern, L, IDs, threads):
    af, pops, its = oracle.reference_af(L, IDs, MAF_ITER, MAF_TOLE, threads, kern)
    ll, _, lits = oracle.loo(L, af.copy(), IDs, threads, MAF_ITER, MAF_TOLE, kern=kern)
    return ll


def cpu_baseline(L, IDs, threads, target_s=12.0, max_sites=None):
    """Time the reference's CPU path on a bounded slice; returns dict for the JSON line."""
    from oracle import oracle
    kind = "reference" if oracle.have_ref() else "port"
    kern = oracle.kernels("ref" if kind == "reference" else "port")
    m_probe = min(L.shape[0], 200)
    t0 = time.perf_counter()
    cpu_step(oracle, kern, np.ascontiguousarray(L[:m_probe]), IDs, threads)
    dt = time.perf_counter() - t0
    m = int(min(L.shape[0], max(m_probe, m_probe * target_s / max(dt, 1e-3))))
    if max_sites:
        m = min(m, max_sites)
    Ls = np.ascontiguousarray(L[:m])
    t0 = time.perf_counter()
    cpu_step(oracle, kern, Ls, IDs, threads)
    dt = time.perf_counter() - t0
    evals = float(m) * N_IND * N_POP
    return {"value": evals / dt, "unit": UNIT, "cores": threads, "kind": kind, "seconds": dt,
            "sample": "first %d sites x %d individuals x %d populations of the same synthetic matrix; "
                      "--get_reference_af + --loo through the restated drivers over %s kernels (-t %d)"
                      % (m, N_IND, N_POP, "the reference's compiled Cython" if kind == "reference" else "the oracle C port", threads)}, m


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle
    from wgsassign_b200 import synth
    threads = os.cpu_count() or 1
    kind = "reference" if oracle.have_ref() else "port"
    kern = oracle.kernels("ref" if kind == "reference" else "port")
    IDs = ids_array()
    # calibrate the slice so that one step is ~8 s of CPU work
    probe = synth.synth(200, N_IND, N_POP, seed=SEED, with_ad=False)["L"]
    t0 = time.perf_counter()
    cpu_step(oracle, kern, probe, IDs, threads)
    dt = time.perf_counter() - t0
    m = int(max(200, min(20000, 200 * 8.0 / max(dt, 1e-3))))
    L = synth.synth(m, N_IND, N_POP, seed=SEED, with_ad=False)["L"]
    for _ in range(args.warmup):
        cpu_step(oracle, kern, L, IDs, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(oracle, kern, L, IDs, threads)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    val = float(m) * N_IND * N_POP / dt
    sample = ("%d sites x %d individuals x %d populations (NumPy generator, same model as the GPU arm); "
              "--get_reference_af + --loo per step through the restated drivers over %s kernels, -t %d"
              % (m, N_IND, N_POP, "the reference's compiled Cython" if kind == "reference" else "the oracle C port", threads))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg3: synthetic 1M sites x 500 individuals x 10 populations, --get_reference_af + --loo "
                                   "(bounded site slice on the CPU)", "sample_sites": m, "individuals": N_IND, "populations": N_POP,
                       "maf_iter": MAF_ITER, "maf_tole": MAF_TOLE},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------
# GPU leg
# ------------------------------------------------------------------------------------------
def ncu_traffic(name, sites):
    """DRAM bytes per launch of a kernel family from the committed ncu capture, scaled to `sites`."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic_r1.json")
    try:
        k = json.load(open(p))["kernels"][name]
        return k["dram_bytes_per_launch"] * sites / k["sites"]
    except Exception:
        return None


def family(ctx, name, peak_gbs, sites=None):
    t = ctx.timing_get(name)
    if t["launches"] == 0 or t["ms"] <= 0:
        return None
    sec = t["ms"] * 1e-3
    out = {"launches": t["launches"], "ms_total": t["ms"], "ms_per_launch": t["ms"] / t["launches"],
           "algorithmic_gb": t["bytes"] / 1e9, "achieved_gbs": t["bytes"] / 1e9 / sec, "hbm_frac": t["bytes"] / 1e9 / sec / peak_gbs,
           "units": t["units"], "units_per_s": t["units"] / sec}
    if sites:
        tr = ncu_traffic(name, sites)
        out["algorithmic_bytes_per_launch"] = t["bytes"] / t["launches"]
        out["ncu_dram_bytes_per_launch"] = tr
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sites", type=int, default=SITES_PER_GPU, help="sites per GPU (default: cfg3's 1M)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the per-kernel extras (pop_like / fisher)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200" and not os.environ.get("WGS_BENCH_ALLOW_SHORT"):
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    from wgsassign_b200 import _lib, dist
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
    M_local = args.sites
    M_total = M_local * world
    if world > 1:
        dist.enable(M_total, rank * M_local, device=torch.device("cuda", local))

    hbm_peak, peak_src, sm_max = peaks()
    ctx = _lib.Context(local)
    pop_of = pop_assignment()
    ctx.set_pops(pop_of, N_POP)
    dist.attach(ctx)           # must precede synth: the generator is keyed by global site index
    ctx.synth(M_local, N_IND, seed=SEED)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as td
            td.barrier()
        torch.cuda.synchronize()

    def step(host_af=False):
        """One pass of --get_reference_af + --loo.  Resident form: the AF matrix stays on the
        device between the two operators; host form (e2e): it is returned to the host, as the
        CLI needs it for <out>.pop_af.npy, and passed back in."""
        af, its = ctx.ref_af(MAF_ITER, MAF_TOLE, download=host_af)
        ll, _, lits = ctx.loo_partial(af, MAF_ITER, MAF_TOLE)
        dist.allreduce_sum(ll)
        return af, ll, its, lits

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.timing_reset(True)
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        af, ll, its, lits = step()
    barrier()
    dt = time.perf_counter() - t0
    launches = ctx.launch_count() - l0
    fam = {k: family(ctx, k, hbm_peak, M_local) for k in ("loo_em", "em_pop", "loo_like", "loo_pack")}
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
    assign_ok = float(np.mean(np.argmax(ll, 1) == pop_of))

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        Lh = _lib.pinned_empty((M_local, 2 * N_IND), np.float32)
        chunk = 100_000
        for s0 in range(0, M_local, chunk):
            n = min(chunk, M_local - s0)
            Lh[s0:s0 + n] = ctx.download(s0, n)
        ctx2 = ctx
        af_pin = _lib.pinned_empty((M_local, N_POP), np.float32)      # host buffer the allele frequencies come back into

        def e2e_step():
            """What the CLI does for `--get_reference_af --loo` on a freshly parsed matrix: queue the
            upload (one strided DMA per population slab), then ONE fused call whose leave-one-out EM
            starts on population 0 while the later slabs are still crossing PCIe; the allele
            frequencies (for <out>.pop_af.npy) and the likelihoods come back to the host."""
            ctx2.set_pops(pop_of, N_POP)
            dist.attach(ctx2)
            if os.environ.get("WGS_E2E_SERIAL"):            # the unpipelined form, for comparison
                ctx2.upload_gl(Lh)
                return step(host_af=True)
            ctx2.upload_gl_async(Lh)
            af_, its_, ll_, _, lits_, _ = ctx2.ref_af_loo(MAF_ITER, MAF_TOLE, af_out=af_pin)
            dist.allreduce_sum(ll_)
            return af_, ll_, its_, lits_
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            af2, ll2, _, _ = e2e_step()
        barrier()
        dte = time.perf_counter() - t0
        if world > 1:
            import torch.distributed as td
            tt = torch.tensor([dte], dtype=torch.float64, device="cuda")
            td.all_reduce(tt, op=td.ReduceOp.MAX)
            dte = float(tt.item())
        serial = bool(os.environ.get("WGS_E2E_SERIAL"))
        h2d = Lh.nbytes + pop_of.nbytes + (af2.nbytes if serial else 0)
        d2h = (2 if serial else 1) * af2.nbytes + 2 * ll2.nbytes + 4 * (N_IND + N_POP)
        e2e = {"value": evals_step / (dte / n_e2e), "unit": UNIT, "h2d_bytes_per_step": int(h2d * world),
               "d2h_bytes_per_step": int(d2h * world), "ms_per_step": dte / n_e2e * 1e3, "steps": n_e2e,
               "identical_to_resident": bool(np.array_equal(ll2, ll)),
               "path": "wgs_upload_gl + wgs_ref_af + wgs_loo_partial" if serial else
                       "wgs_upload_gl_async + wgs_ref_af_loo (leave-one-out EM of population k overlaps the upload of k+1..)"}

    # ---- extras: the other kernels of the path on the same resident matrix (untimed for `value`) ----
    extra = {}
    if af is None:
        af, _ = ctx.ref_af(MAF_ITER, MAF_TOLE)
    if not args.no_extra:
        ctx.pop_like_partial(af); ctx.fisher_partial(af)        # untimed first use (module load, pool growth)
        ctx.timing_reset(True)
        for _ in range(3):
            pl = ctx.pop_like_partial(af)
            fo = ctx.fisher_partial(af)
        extra = {k: family(ctx, k, hbm_peak, M_local) for k in ("pop_like", "fisher")}
        aux = ctx.timing_get("pop_like_aux")                    # coefficient rows + population-only log term of the ratio form
        if extra.get("pop_like") and aux["launches"]:
            pl_ = extra["pop_like"]
            pl_["aux_ms_per_call"] = aux["ms"] / pl_["launches"]
            pl_["ms_per_call_with_aux"] = (pl_["ms_total"] + aux["ms"]) / pl_["launches"]
            pl_["hbm_frac_with_aux"] = pl_["algorithmic_gb"] / ((pl_["ms_total"] + aux["ms"]) * 1e-3) / hbm_peak
        for mode, nm in ((0, "stream_flat"), (1, "stream_slab")):       # what a pure read of the same matrix achieves
            ms, nb = ctx.debug_stream(mode)
            extra[nm] = {"ms": ms, "gb": nb / 1e9, "achieved_gbs": nb / 1e9 / (ms * 1e-3)}
        ctx.timing_reset(False)

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
    # The packed kernel (loo_em_step5) spends 8 FFMA2 and 2 reciprocals per 4 posterior evaluations, so the
    # FP32 pipe binds (2.0 packed lane-instructions per evaluation), MUFU sits at 0.5 per evaluation.
    fp32_peak = 148 * 58.0 * sm_mhz * 1e6 / 2.0
    mufu_peak = 148 * 15.48 * sm_mhz * 1e6 / 0.5
    le = fam["loo_em"]
    roofline = {"kernel": "loo_em_step5_kernel", "bound": "hbm", "achieved": le["achieved_gbs"], "peak": hbm_peak,
                "unit": "GB/s", "frac": le["hbm_frac"], "traffic": le.get("ncu_dram_bytes_per_launch"),
                "algorithmic_bytes_per_launch": le.get("algorithmic_bytes_per_launch"), "peak_source": peak_src,
                "ms_per_launch": le["ms_per_launch"], "launches": le["launches"],
                "share_of_step": le["ms_total"] / (dt * 1e3),
                "note": "by design NOT HBM-bound: each packed row group is read once per iteration and re-used for n^2 posterior "
                        "evaluations from shared memory; the binding limit is the packed-FP32 issue rate (see `issue`); "
                        "the HBM-bound kernels of the path are in `kernels` (em_pop, pop_like, fisher)",
                "issue": {"bound": "fp32 pipe (FFMA2)", "achieved": le["units_per_s"], "peak": fp32_peak,
                          "peak_source": "measured 58.0 packed FP32 lane-instructions/clk/SM (profiles/issue_rates_r1.txt), "
                                         "2.0 per posterior evaluation",
                          "unit": "posterior evals/s", "frac": le["units_per_s"] / fp32_peak, "sm_mhz": sm_mhz,
                          "mufu_rcp_frac": le["units_per_s"] / mufu_peak,
                          "inner_loop_ceiling": "6.3e12 evals/s for the bare inner loop from shared memory "
                                                "(scripts/microbench/loo_quad_rate.cu, profiles/loo_quad_rate_r1.txt)"}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "cfg3: synthetic 1M sites x 500 individuals x 10 populations per GPU, "
                                   "--get_reference_af + --loo (BASELINE.json configs[2])",
                       "sites_per_gpu": M_local, "sites_total": M_total, "individuals": N_IND, "populations": N_POP,
                       "maf_iter": MAF_ITER, "maf_tole": MAF_TOLE, "parallelism": "site-sharded x%d" % world,
                       "l2": "inputs (%.1f GB GL per GPU) exceed the 126 MB L2, no flush needed" % (M_local * N_IND * 8 / 1e9),
                       "timer": "host clock around blocking C-ABI calls, device-synchronised + barrier on both sides, max over ranks; "
                                "kernels timed with CUDA events on the library's stream",
                       "em_iters_ref": [int(x) for x in its], "em_iters_loo_minmax": [int(np.min(lits)), int(np.max(lits))],
                       "self_assignment_rate": assign_ok},
            "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": roofline,
            "kernels": {**fam, **extra}}
    if not args.no_cpu_baseline and world == 1:
        m_dl = min(M_local, 20000)
        Ls = ctx.download(0, m_dl)
        cb, m_used = cpu_baseline(Ls, ids_array(), os.cpu_count() or 1)
        line["cpu_baseline"] = cb
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
