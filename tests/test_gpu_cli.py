"""End-to-end: the drop-in CLI (same flags, same output files) on the GPU against the files
the unmodified reference CLI wrote for the same inputs (tests/golden)."""
import os

import numpy as np
import pytest

from conftest import parse_tsv

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cli():
    import wgsassign_b200._lib as _lib
    if _lib.lib().wgs_device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box")
    from wgsassign_b200 import WGSassign, session
    session.reset()
    yield WGSassign
    session.reset()


def write_inputs(tmp, L, samples, sites, IDs):
    from wgsassign_b200 import synth
    bg = os.path.join(tmp, "in.beagle.gz")
    synth.write_beagle(bg, L, list(samples), list(sites))
    ids = os.path.join(tmp, "ids.txt")
    np.savetxt(ids, IDs, fmt="%s", delimiter="\t")
    return bg, ids


def rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30))


def test_cli_config1_and_2(cli, bundled, tmp_path, capsys):
    tmp = str(tmp_path)
    bg, ids = write_inputs(tmp, bundled["L_breeding"], bundled["samples_breeding"], bundled["sites_breeding"], bundled["IDs_breeding"])
    out = os.path.join(tmp, "c1")
    cli.main(["--beagle", bg, "--pop_af_IDs", ids, "--get_reference_af", "--ne_obs", "--loo", "--out", out, "-t", "4"])
    txt = capsys.readouterr().out
    assert "Loaded 449 sites and 85 individuals." in txt and "Performing leave-one-out cross validation." in txt
    assert txt.count("EM (MAF) converged at iteration:") == 5 + 85
    af = np.load(out + ".pop_af.npy")
    assert af.dtype == np.float32 and np.max(np.abs(af - bundled["c1_pop_af"])) < 1e-5
    assert list(np.loadtxt(out + ".pop_names.txt", dtype=str)) == list(bundled["c1_pop_names"])
    hdr, rows = parse_tsv(open(out + ".pop_like_LOO.tsv").read())
    ghdr, grows = parse_tsv(str(bundled["c1_loo_tsv"]))
    assert hdr == ghdr and [r[:2] for r in rows] == [r[:2] for r in grows]
    got = np.array([[float(x) for x in r[2:]] for r in rows]); gold = np.array([[float(x) for x in r[2:]] for r in grows])
    assert rel(got, gold) < 2e-6 and np.array_equal(np.argmax(got, 1), np.argmax(gold, 1))
    gne = np.array([float(x) for x in str(bundled["c1_ne_obs_txt"]).split("\n")[1].split()])
    ne = np.array([float(x) for x in open(out + ".ne_obs.txt").read().split("\n")[1].split()])
    assert rel(ne, gne) < 1e-5
    gind = np.array([float(x) for x in str(bundled["c1_ne_ind_txt"]).split()])
    assert rel(np.loadtxt(out + ".ne_ind.txt"), gind) < 1e-5
    assert np.load(out + ".fisher_obs.npy").shape == (449, 5) and np.load(out + ".ne_obs.npy").shape == (449, 5)
    assert "-get_reference_af" in open(out + ".args").read()
    # config 2: assign the non-breeding birds with the AF file just written
    bg2, _ = write_inputs(tmp, bundled["L_nonbreeding"], bundled["samples_nonbreeding"], bundled["sites_breeding"], bundled["IDs_nonbreeding"])
    out2 = os.path.join(tmp, "c2")
    cli.main(["--beagle", bg2, "--pop_af_file", out + ".pop_af.npy", "--get_pop_like", "--out", out2])
    pl = np.loadtxt(out2 + ".pop_like.txt")
    assert rel(pl, bundled["c2_pop_like"]) < 2e-6
    assert np.array_equal(np.argmax(pl, 1), np.argmax(bundled["c2_pop_like"], 1))


def test_cli_downsampled_partitions(cli, bundled, tmp_path):
    tmp = str(tmp_path)
    bg, ids = write_inputs(tmp, bundled["L_breeding"], bundled["samples_breeding"], bundled["sites_breeding"], bundled["IDs_breeding"])
    from wgsassign_b200 import synth
    ds = os.path.join(tmp, "ds.beagle.gz")
    synth.write_beagle(ds, bundled["L_breeding_ds"], list(bundled["samples_breeding"]), list(bundled["sites_breeding_ds"]))
    out = os.path.join(tmp, "d")
    cli.main(["--beagle", bg, "--pop_af_IDs", ids, "--get_reference_af", "--loo", "--loo_downsampled_beagle", ds,
              "--partition_sites", "3", "--out", out])
    assert np.load(out + ".pop_af.npy").shape == (357, 5)         # filtered BEFORE the AF estimate (WGSassign.py:189)
    _, rows = parse_tsv(open(out + ".pop_like_LOO_downsampled.tsv").read())
    _, grows = parse_tsv(str(bundled["c1ds_loo_tsv"]))
    got = np.array([[float(x) for x in r[2:]] for r in rows]); gold = np.array([[float(x) for x in r[2:]] for r in grows])
    assert rel(got, gold) < 2e-6
    import gzip
    hdr, prow = parse_tsv(gzip.open(out + ".pop_like_LOO_downsampled_partitions_3.tsv.gz", "rt").read())
    ghdr, gprow = parse_tsv(str(bundled["c1ds_loo_parts_tsv"]))
    assert hdr == ghdr and [r[:3] for r in prow] == [r[:3] for r in gprow]
    got = np.array([[float(x) for x in r[3:]] for r in prow]); gold = np.array([[float(x) for x in r[3:]] for r in gprow])
    assert rel(got, gold) < 5e-6


def test_cli_zscores(cli, zgold, tmp_path, capsys):
    tmp = str(tmp_path)
    L, AD, IDs, af = zgold["L"], zgold["AD"].astype(np.int32), zgold["IDs"], zgold["af"]
    sites = ["chr1_%d" % (100 + 7 * s) for s in range(L.shape[0])]
    bg, ids = write_inputs(tmp, L, IDs[:, 0], sites, IDs)
    adf = os.path.join(tmp, "ad.txt")
    np.savetxt(adf, AD, fmt="%d")
    afp = os.path.join(tmp, "af.npy")
    np.save(afp, af)
    pn = os.path.join(tmp, "pops.txt")
    np.savetxt(pn, np.unique(IDs[:, 1]), fmt="%s")
    out = os.path.join(tmp, "z")
    cli.main(["--beagle", bg, "--pop_af_IDs", ids, "--pop_af_file", afp, "--pop_names", pn, "--ind_ad_file", adf,
              "--get_assignment_z_score", "--out", out])
    txt = capsys.readouterr().out
    assert [int(x.split(": ")[1]) for x in txt.split("\n") if x.startswith("Loci used")] == list(zgold["z_assign_loci"])
    z = np.loadtxt(out + ".z_ind.txt")
    gz = np.array([float(x) for x in str(zgold["z_assign_txt"]).split()])
    # z = (W_obs - z_mu) / sqrt(z_var) subtracts two sums of magnitude ~|W|: the 1e-6 relative tolerance of the components
    # propagates to 2e-6 (|W_obs| + |z_mu|) / sqrt(z_var) on z (hard part 4) - at most 4e-4 on this fixture
    comp = zgold["z_assign_components"]
    tol = 2e-6 * (np.abs(comp[:, 0]) + np.abs(comp[:, 1])) / np.sqrt(comp[:, 2]) + 1e-5 * np.abs(gz)
    assert np.all(np.abs(z - gz) <= tol) and z.shape == (12,)
    cli.main(["--beagle", bg, "--pop_af_IDs", ids, "--pop_names", pn, "--ind_ad_file", adf, "--get_reference_z_score",
              "--ind_start", "2", "--ind_end", "9", "--out", out])
    z = np.loadtxt(out + ".reference_z_ind.txt")
    gz = np.array([float(x) for x in str(zgold["z_ref_txt"]).split()])[2:9]
    assert np.max(np.abs(z - gz)) < 2e-4 and z.shape == (7,)
    with pytest.raises(AssertionError):
        cli.main(["--beagle", bg, "--pop_af_IDs", ids, "--pop_names", pn, "--ind_ad_file", adf, "--get_reference_z_score",
                  "--ind_start", "0", "--out", out])
