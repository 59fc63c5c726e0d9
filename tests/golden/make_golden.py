"""Generates tests/golden/*.npz by running the UNMODIFIED reference in this container.

Needs /root/reference (Python drivers + bundled data) and oracle/_ref (the reference's
compiled Cython kernels, from oracle/build_ref.sh).  The GPU box has neither the
reference checkout nor a way to regenerate these, so the outputs are committed.

    python tests/golden/make_golden.py

What is recorded:
  bundled.npz   configs 1-2 of BASELINE.json on the bundled data: parsed GL matrices (the
                reader's output, i.e. the input contract), --get_reference_af, --ne_obs,
                --loo (plain, and downsampled + --partition_sites 3), --get_pop_like,
                --get_em_mix; all through the reference CLI, outputs read back from files.
  zscore.npz    seeded synthetic GL + allele depths (the reference bundles no AD file):
                --get_assignment_z_score and --get_reference_z_score through the CLI plus
                per-individual components from the reference's zscore.py functions.
"""
import contextlib
import io
import math
import os
import re
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("WGS_REFERENCE_DIR", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, ROOT)
np.math = math  # zscore.py:73 uses np.math.factorial, removed from numpy >= 1.25

from WGSassign import WGSassign as ref_cli  # noqa: E402
from WGSassign import reader_cy, zscore as ref_zscore  # noqa: E402
from wgsassign_b200 import synth  # noqa: E402

DATA = os.path.join(REF, "data")
BREED = os.path.join(DATA, "amre.breeding.ind85.ds_2x.sites-filter.top_50_each.beagle.gz")
BREED_DS = os.path.join(DATA, "amre.breeding.ind85.ds_2x.sites-filter.top_50_each_subset_80percent_sites.beagle.gz")
BREED_IDS = os.path.join(DATA, "amre.breeding.ind85.reference_k5.IDs.txt")
NONB = os.path.join(DATA, "amre.nonbreeding.ind34.ds_2x.sites-filter.top_50_each.beagle.gz")
NONB_IDS = os.path.join(DATA, "amre.nonbreeding.ind34.site.IDs.txt")


def run_cli(argv):
    """Run the reference's main() with argv; returns captured stdout."""
    old = sys.argv
    sys.argv = ["WGSassign"] + argv
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            ref_cli.main()
    finally:
        sys.argv = old
    return buf.getvalue()


def em_iters(stdout):
    return np.array([int(x) for x in re.findall(r"EM \(MAF\) converged at iteration: (\d+)", stdout)], np.int32)


def read_tsv(path):
    import gzip
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as fh:
        return fh.read()


def bundled(tmp):
    out = {}
    L, samples, sites = reader_cy.readBeagle(BREED)
    L_ds, samples_ds, sites_ds = reader_cy.readBeagle(BREED_DS)
    L_nb, samples_nb, sites_nb = reader_cy.readBeagle(NONB)
    out.update(L_breeding=L, samples_breeding=np.array(samples), sites_breeding=np.array(sites),
               L_breeding_ds=L_ds, sites_breeding_ds=np.array(sites_ds),
               L_nonbreeding=L_nb, samples_nonbreeding=np.array(samples_nb),
               IDs_breeding=np.loadtxt(BREED_IDS, delimiter="\t", dtype=str),
               IDs_nonbreeding=np.loadtxt(NONB_IDS, delimiter="\t", dtype=str))

    # config 1: --get_reference_af --ne_obs --loo
    p = os.path.join(tmp, "c1")
    so = run_cli(["--beagle", BREED, "--pop_af_IDs", BREED_IDS, "--get_reference_af", "--ne_obs",
                  "--loo", "--out", p])
    its = em_iters(so)
    K = 5
    out.update(c1_pop_af=np.load(p + ".pop_af.npy"), c1_pop_names=np.loadtxt(p + ".pop_names.txt", dtype=str),
               c1_em_iters_ref=its[:K], c1_em_iters_loo=its[K:],
               c1_fisher_obs=np.load(p + ".fisher_obs.npy"), c1_ne_obs=np.load(p + ".ne_obs.npy"),
               c1_ne_obs_txt=read_tsv(p + ".ne_obs.txt"), c1_ne_ind_txt=read_tsv(p + ".ne_ind.txt"),
               c1_loo_tsv=read_tsv(p + ".pop_like_LOO.tsv"))
    # config 1 with downsampled LOO + partitions
    p2 = os.path.join(tmp, "c1ds")
    so = run_cli(["--beagle", BREED, "--pop_af_IDs", BREED_IDS, "--get_reference_af", "--loo",
                  "--loo_downsampled_beagle", BREED_DS, "--partition_sites", "3", "--out", p2])
    its = em_iters(so)
    out.update(c1ds_pop_af=np.load(p2 + ".pop_af.npy"), c1ds_em_iters_loo=its[K:],
               c1ds_loo_tsv=read_tsv(p2 + ".pop_like_LOO_downsampled.tsv"),
               c1ds_loo_parts_tsv=read_tsv(p2 + ".pop_like_LOO_downsampled_partitions_3.tsv.gz"))
    # config 2: --get_pop_like (+ mixture)
    p3 = os.path.join(tmp, "c2")
    run_cli(["--beagle", NONB, "--pop_af_file", p + ".pop_af.npy", "--get_pop_like", "--out", p3])
    out.update(c2_pop_like_txt=read_tsv(p3 + ".pop_like.txt"), c2_pop_like=np.loadtxt(p3 + ".pop_like.txt"))
    run_cli(["--pop_like", p3 + ".pop_like.txt", "--pop_like_IDs", NONB_IDS, "--get_em_mix", "--out", p3])
    out.update(c2_em_mix_txt=read_tsv(p3 + ".em_mix.txt"))
    np.savez_compressed(os.path.join(HERE, "bundled.npz"), **out)
    print("bundled.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})


def zscore_fixture(tmp):
    """Synthetic depth-consistent data; the reference CLI reads everything from files."""
    m, n, k = 3000, 12, 3
    d = synth.synth(m, n, k, seed=20261018, depth=2.0)
    L, AD, IDs = d["L"], d["AD"], d["IDs"]
    sites = ["chr1_%d" % (100 + 7 * s) for s in range(m)]
    bg = os.path.join(tmp, "z.beagle.gz")
    synth.write_beagle(bg, L, list(IDs[:, 0]), sites)
    Lr, _, _ = reader_cy.readBeagle(bg)
    assert np.array_equal(Lr, L), "beagle text round trip must reproduce the float32 matrix"
    ids = os.path.join(tmp, "z.IDs.txt")
    np.savetxt(ids, IDs, fmt="%s", delimiter="\t")
    adf = os.path.join(tmp, "z.ad.txt")
    np.savetxt(adf, AD, fmt="%d")
    p = os.path.join(tmp, "z")
    so = run_cli(["--beagle", bg, "--pop_af_IDs", ids, "--get_reference_af", "--out", p])
    af = np.load(p + ".pop_af.npy")
    out = dict(L=L, AD=AD.astype(np.int16), IDs=IDs, af=af, em_iters_ref=em_iters(so))
    so = run_cli(["--beagle", bg, "--pop_af_IDs", ids, "--pop_af_file", p + ".pop_af.npy",
                  "--pop_names", p + ".pop_names.txt", "--ind_ad_file", adf,
                  "--get_assignment_z_score", "--out", p])
    out["z_assign_txt"] = read_tsv(p + ".z_ind.txt")
    out["z_assign_loci"] = np.array([int(x) for x in re.findall(r"Loci used: (\d+)", so)], np.int64)
    so = run_cli(["--beagle", bg, "--pop_af_IDs", ids, "--pop_names", p + ".pop_names.txt",
                  "--ind_ad_file", adf, "--get_reference_z_score", "--out", p])
    out["z_ref_txt"] = read_tsv(p + ".reference_z_ind.txt")
    out["z_ref_loci"] = np.array([int(x) for x in re.findall(r"Loci used: (\d+)", so)], np.int64)
    out["z_ref_em_iters"] = em_iters(so)
    # threshold variants through the CLI
    so = run_cli(["--beagle", bg, "--pop_af_IDs", ids, "--pop_af_file", p + ".pop_af.npy",
                  "--pop_names", p + ".pop_names.txt", "--ind_ad_file", adf, "--get_assignment_z_score",
                  "--allele_count_threshold", "25", "--ind_start", "2", "--ind_end", "9", "--out", p + "t"])
    out["z_assign_thr25_txt"] = read_tsv(p + "t.z_ind.txt")
    out["z_assign_thr25_loci"] = np.array([int(x) for x in re.findall(r"Loci used: (\d+)", so)], np.int64)
    so = run_cli(["--beagle", bg, "--pop_af_IDs", ids, "--pop_af_file", p + ".pop_af.npy",
                  "--pop_names", p + ".pop_names.txt", "--ind_ad_file", adf, "--get_assignment_z_score",
                  "--single_read_threshold", "--out", p + "s"])
    out["z_assign_single_txt"] = read_tsv(p + "s.z_ind.txt")
    out["z_assign_single_loci"] = np.array([int(x) for x in re.findall(r"Loci used: (\d+)", so)], np.int64)
    # per-individual components straight from the reference's zscore.py (assignment mode)
    pops = np.loadtxt(p + ".pop_names.txt", dtype=str)
    comps = []
    ADi = AD.astype(np.int32)
    for i in range(n):
        kk = int(np.argwhere(pops == IDs[i, 1])[0][0])
        dct, arr = ref_zscore.AD_summary(L, ADi, i, 0, False)
        keep, kept = ref_zscore.get_L_keep(L, ADi, dct, arr, i)
        fac, like, idx = ref_zscore.get_factorials(arr, dct, 0.01)
        afv = np.ascontiguousarray(af[keep, :][:, kk].reshape(-1))
        w_obs, w_l = ref_zscore.get_expected_W_l(L, keep, afv, ADi, arr, fac, like, idx, 1, i)
        var = ref_zscore.get_var_W_l(L, keep, afv, ADi, arr, fac, like, idx, w_l, 1, i)
        comps.append([w_obs, np.sum(w_l), np.sum(var), kept])
        out["AD_array_%d" % i] = arr
        out["L_keep_%d" % i] = keep
    out["z_assign_components"] = np.array(comps, np.float64)
    np.savez_compressed(os.path.join(HERE, "zscore.npz"), **out)
    print("zscore.npz:", {k: getattr(v, "shape", None) for k, v in out.items() if not k.startswith(("AD_array_", "L_keep_"))})
    print(out["z_assign_txt"].split()[:6], out["z_ref_txt"].split()[:6])


if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as tmp:
        bundled(tmp)
        zscore_fixture(tmp)
