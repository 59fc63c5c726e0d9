"""Command-line front end with the reference's surface (WGSassign/WGSassign.py:24-104 flags,
:109-472 behaviour): same flags, same output files and formats, same progress messages.
The per-population / per-individual Python loops of the reference are replaced by one call
per mode into the GPU library.

    python -m wgsassign_b200.WGSassign --beagle x.beagle.gz --pop_af_IDs ids.txt --get_reference_af --loo

Under `torchrun` every rank parses the file, keeps its contiguous site range on its own GPU
and rank 0 writes the outputs.
"""
import argparse
import os
import sys
from datetime import datetime


def build_parser():
    p = argparse.ArgumentParser(prog="WGSassign")
    p.add_argument("-b", "--beagle", metavar="FILE", help="Filepath to genotype likelihoods in gzipped Beagle format from ANGSD")
    p.add_argument("-t", "--threads", metavar="INT", type=int, default=1, help="Number of threads")
    p.add_argument("-o", "--out", metavar="OUTPUT", default="wgsassign", help="Prefix for output files")
    p.add_argument("--maf_iter", metavar="INT", type=int, default=200,
                   help="Maximum iterations for minor allele frequencies estimation - EM (200)")
    p.add_argument("--maf_tole", metavar="FLOAT", type=float, default=1e-4,
                   help="Tolerance for minor allele frequencies estimation update - EM (1e-4)")
    p.add_argument("--pop_af_IDs", metavar="FILE", help="Filepath to individual IDs and populations for beagle")
    p.add_argument("--get_reference_af", action="store_true", help="Estimate allele frequencies for reference populations")
    p.add_argument("--pop_names", metavar="FILE", help="Filepath to population names of allele frequency file")
    p.add_argument("--ne_obs", action="store_true", help="Estimate population and individuals effective sample sizes")
    p.add_argument("--loo", action="store_true", help="Perform leave-one-out cross validation")
    p.add_argument("--loo_downsampled_beagle", metavar="FILE",
                   help="Optional Beagle file of downsampled genotype likelihoods to use for LOO assignment. "
                        "(To test accuracy when assigned samples have lower coverage)")
    p.add_argument("--pop_af_file", metavar="FILE", help="Filepath to reference population allele frequencies")
    p.add_argument("--get_pop_like", action="store_true",
                   help="Estimate log likelihood of individual assignment to each reference population")
    p.add_argument("--partition_sites", type=int, metavar="INT", default=1,
                   help="Optional: partition sites into INT subsets (by modulo) and report assignment log-likelihoods for each subset.")
    p.add_argument("--get_assignment_z_score", action="store_true", help="Calculate z-score for individuals")
    p.add_argument("--get_reference_z_score", action="store_true", help="Calculate z-score for individuals")
    p.add_argument("--ind_ad_file", metavar="FILE", help="Filepath to individual allele depths, tab-delimited, .txt or .gz")
    p.add_argument("--allele_count_threshold", metavar="INT", type=int,
                   help="Minimum number of loci needed to keep a specific allele count combination")
    p.add_argument("--single_read_threshold", action="store_true",
                   help="Use only loci with a single read. Helpful for computational efficiency when individuals's sequencing depths vary.")
    p.add_argument("--ind_start", metavar="INT", type=int,
                   help="Start analysis at this individual index (0-index: i.e. 0 starts with the 1st individual)")
    p.add_argument("--ind_end", metavar="INT", type=int,
                   help="End analysis at this individual index (0-index: i.e. If you have 10 individuals, 9 is the 10th individual)")
    p.add_argument("--pop_like", metavar="FILE", help="Filepath to population assignment log likelihood file")
    p.add_argument("--pop_like_IDs", metavar="FILE", help="Filepath to IDs for population assignment log likelihood file")
    p.add_argument("--get_em_mix", action="store_true", help="Estimate mixture proportions with EM algorithm")
    p.add_argument("--get_mcmc_mix", action="store_true", help="Estimate mixture proportions with MCMC algorithm")
    p.add_argument("--mixture_iter", metavar="INT", type=int, default=200, help="Maximum iterations mixture estimation - EM (200)")
    p.add_argument("--em_mix_logsumexp", action="store_true",
                   help="(not in the reference) subtract each individual's largest log likelihood before exponentiating in --get_em_mix: "
                        "finite mixture proportions for genome-scale log likelihoods, identical results where the reference's are finite")
    p.add_argument("--shard_by_bytes", action="store_true",
                   help="(not in the reference; under torchrun only) when the Beagle file is BGZF (bgzip, as ANGSD writes it) every rank "
                        "reads its own byte range of the file instead of inflating all of it: the ranks then hold unequal numbers of sites; "
                        "results do not depend on how the sites are split")
    return p


parser = build_parser()


def _write_args_file(args):
    """`<out>.args`: time, directory and every non-default option (WGSassign.py:127-141)."""
    chosen, defaults = vars(args), vars(parser.parse_args([]))
    with open(args.out + ".args", "w") as fh:
        fh.write("WGSassign\n")
        fh.write("Time: " + datetime.now().strftime("%d/%m/%Y %H:%M:%S") + "\n")
        fh.write("Directory: " + str(os.getcwd()) + "\n")
        fh.write("Options:\n")
        for key, val in chosen.items():
            if val != defaults[key]:
                fh.write("\t-" + str(key) + ("\n" if isinstance(val, bool) else " " + str(val) + "\n"))


class _Run:
    """State shared by the modes of one invocation."""

    def __init__(self, args):
        import numpy as np
        from . import dist
        self.np, self.args, self.dist = np, args, dist
        self.L = self.L_ds = None
        self.sample_names = self.site_names = None
        self.rank0 = True
        self.M_total = None
        self._init_dist()

    # -- site sharding under torchrun ---------------------------------------------------
    def _init_dist(self):
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            import torch
            import torch.distributed as td
            local = int(os.environ.get("LOCAL_RANK", "0"))
            use_cuda = torch.cuda.is_available()
            if use_cuda:
                torch.cuda.set_device(local)
                self.dist.bind_near_gpu(local)               # pinned read buffers on the GPU's NUMA node
            if not td.is_initialized():
                td.init_process_group("nccl" if use_cuda else "gloo")
            self.rank0 = td.get_rank() == 0
            self._device = torch.device("cuda", local) if use_cuda else None

    def _world(self):
        return int(os.environ.get("WORLD_SIZE", "1"))

    def _range(self, M):
        """This rank's contiguous site range of an M-site input (and the shard geometry for the library)."""
        import torch.distributed as td
        if getattr(self, "_ranges", None) is not None:             # byte-range parts: the ranks' row ranges are what they read
            assert M == self.M_total, "a per-site input has %d rows, the Beagle file %d" % (M, self.M_total)
            return self._ranges[td.get_rank()]
        lo, hi = self.dist.shard_range(M, td.get_rank(), td.get_world_size())
        if self.M_total is None:
            self.M_total = M
            self.dist.enable(self.M_total, lo, device=self._device)
        return lo, hi

    def shard(self, X):
        """Keep this rank's contiguous site range of a per-site matrix that was read whole."""
        if self._world() <= 1:
            return X
        lo, hi = self._range(X.shape[0])
        return self.np.ascontiguousarray(X[lo:hi])

    def full(self, X):
        """Per-site output of this rank -> the full matrix (every rank)."""
        return self.dist.gather_rows(X)

    def say(self, *a):
        if self.rank0:
            print(*a)

    def _first_layout(self):
        """Population layout of the first mode that will use the matrix: (pop_of_ind, K) or (None, 0) for the flat one.
        The streamed upload puts the matrix on the device in that layout while the file is still being parsed."""
        np, a = self.np, self.args
        from . import session
        try:
            if a.get_reference_af or (a.get_reference_z_score and not a.get_pop_like):
                IDs = np.loadtxt(a.pop_af_IDs, delimiter="\t", dtype="str")
                pop_of, pops = session.pops_from_ids(IDs)
                return pop_of, len(pops)
        except Exception:
            pass                                                   # the mode itself reports a missing / malformed ID file
        return None, 0

    # -- input ----------------------------------------------------------------------------
    def parse_inputs(self):
        from . import reader, session, utils
        a = self.args
        if a.beagle is not None:
            self.say("Parsing Beagle file.")
            assert os.path.isfile(a.beagle), "Beagle file doesn't exist!"
            if a.loo_downsampled_beagle is None:
                # streamed: a background thread inflates, a thread pool parses straight into pinned memory, every finished
                # block is queued for upload while the next is parsed; a site-sharded rank converts only its own rows
                rows = part = None
                if self._world() > 1:
                    import torch.distributed as td
                    if a.shard_by_bytes and reader.is_bgzf(a.beagle):
                        part = (td.get_rank(), td.get_world_size())          # every rank inflates only its share of the file
                    else:
                        if a.shard_by_bytes:
                            self.say("--shard_by_bytes: not a BGZF file, every rank reads through the whole file.")
                        m_all, _, _ = reader.count_rows(a.beagle, a.threads)      # one inflate pass, nothing converted
                        rows = self._range(m_all)
                pop_of, K = self._first_layout()
                ctx, self.L, self.sample_names, self.site_names = session.stream_context(a.beagle, pop_of, K, a.threads, rows=rows, part=part)
                self._L_sharded = True
                m, n = len(self.site_names), self.L.shape[1] // 2
                if part is not None:
                    # the ranks hold consecutive, unequal pieces: their row counts give the geometry (the site names of this
                    # rank are its own rows only)
                    self.M_total, lo, self._ranges = self.dist.row_counts_to_ranges(self.L.shape[0])
                    self.dist.enable(self.M_total, lo, device=self._device, ranges=self._ranges)
                    self.dist.attach(ctx)
                    m = self.M_total
            else:
                self.L, self.sample_names, self.site_names = reader.readBeagle(a.beagle, a.threads)
                m, n = self.L.shape[0], self.L.shape[1] // 2
            self.say("Loaded " + str(m) + " sites and " + str(n) + " individuals.")
            if self.rank0:
                utils.print_sample_and_site_summary(self.sample_names, self.site_names)
        if a.loo_downsampled_beagle is not None:
            self.say("Parsing the optional downsampled Beagle file.")
            assert os.path.isfile(a.loo_downsampled_beagle), "Downsampled beagle file doesn't exist!"
            L_ds, names_ds, sites_ds = reader.readBeagle(a.loo_downsampled_beagle, a.threads)
            # the reference reports the ORIGINAL shape here (WGSassign.py:177-179)
            self.say("Loaded optional downsampled data set with " + str(self.L.shape[0]) + " sites and "
                     + str(self.L.shape[1] // 2) + " individuals.")
            if self.rank0:
                utils.print_sample_and_site_summary(names_ds, sites_ds)
            if self.sample_names != names_ds:
                raise ValueError("Sample names in downsampled Beagle file do not match original.")
            self.say("Retaining only sites from the reference that are in the downsampled beagle file:")
            self.L, self.site_names = utils.filter_sites_to_common(self.L, self.site_names, sites_ds)
            self.say("Removing sites from downsampled set that were not in the reference (should not occur...):")
            L_ds, sites_ds = utils.filter_sites_to_common(L_ds, sites_ds, self.site_names)
            if self.site_names != sites_ds:
                raise ValueError("Site names in full and downsampled Beagle do not match after filtering.")
            self.L_ds = self.shard(self.np.ascontiguousarray(L_ds))
        if self.L is not None and not getattr(self, "_L_sharded", False):
            self.L = self.shard(self.np.ascontiguousarray(self.L))

    # -- --get_reference_af (+ --ne_obs, --loo) ------------------------------------------------
    def reference_af(self):
        np, a = self.np, self.args
        from . import fisher, glassy, session, utils
        self.say("Parsing reference population ID file.")
        assert os.path.isfile(a.pop_af_IDs), "Reference population ID file does not exist!!"
        IDs = np.loadtxt(a.pop_af_IDs, delimiter="\t", dtype="str")
        pops = np.unique(IDs[:, 1])
        assert (self.L.shape[1] // 2 == IDs.shape[0]), "Number of individuals in beagle and reference ID file do not match!"
        pop_of, _ = session.pops_from_ids(IDs)
        # with --loo both operators run as one fused device call whose leave-one-out EM overlaps the upload
        ctx = session.context(self.L, pop_of, len(pops), async_upload=bool(a.loo))
        loo_out = None
        if a.loo:
            if self.L_ds is not None:
                session.with_downsampled(ctx, self.L_ds)
            af, iters, ll, ll_parts, loo_iters = glassy.loo_fused(ctx, IDs, a.maf_iter, a.maf_tole,
                                                                  downsampled=self.L_ds is not None,
                                                                  num_partitions=a.partition_sites)
            loo_out = (ll, ll_parts, loo_iters)
        else:
            af, iters = ctx.ref_af(a.maf_iter, a.maf_tole)
        for it in iters:
            if it > 0:
                self.say("EM (MAF) converged at iteration: " + str(int(it)))
        af_full = self.full(af)
        if self.rank0:
            np.save(a.out + ".pop_af", af_full)
            print("Saved reference population allele frequencies as " + str(a.out) + ".pop_af.npy (Binary - np.float32)\n")
            print("Column order of populations is: " + str(pops))
            np.savetxt(a.out + ".pop_names.txt", pops, fmt="%s")
            print("Saved reference population names as " + str(a.out)
                  + ".pop_names.txt (String: Order of pops for .pop_af.npy, .ne_obs.npy, and fisher_obs.npy files)\n")
        if a.ne_obs:
            self.say("Estimating Fisher information.")
            f_obs, ne_obs = fisher.fisher_obs(self.L, af, IDs, a.threads)
            f_obs, ne_obs = self.full(f_obs), self.full(ne_obs)
            self.say("Estimating individual effective sample sizes.")
            ne_ind = fisher.fisher_obs_ind(self.L, af, IDs, a.threads)
            if self.rank0:
                np.save(a.out + ".fisher_obs", f_obs)
                print("Saved reference population observed Fisher information per locus as " + str(a.out)
                      + ".fisher_obs.npy (Binary - np.float32)\n")
                np.save(a.out + ".ne_obs", ne_obs)
                print("Saved reference population effective sample size estimates per locus as " + str(a.out)
                      + ".ne_obs.npy (Binary - np.float32)\n")
                table = np.empty((2, len(pops)), dtype=np.dtype("U25"))
                table[0, :] = pops
                table[1, :] = np.mean(ne_obs, axis=0)
                np.savetxt(a.out + ".ne_obs.txt", table, fmt="%s")
                print("Saved reference population effective sample size estimates as " + str(a.out) + ".ne_obs.txt (String - np.U25)\n")
                np.savetxt(a.out + ".ne_ind.txt", ne_ind.reshape(-1, 1), fmt="%.7f")
                print("Save individual effective sample sizes as " + str(a.out) + ".ne_ind.txt")
        if a.loo:
            self.say("Performing leave-one-out cross validation.")
            ll, ll_parts, loo_iters = loo_out
            self.say(str(IDs.shape[0]) + " individuals to assign to " + str(len(pops)) + " populations")
            if self.L_ds is not None:
                self.say("Using downsampled GLs for likelihood evaluation in LOO assignment.")
            for it in loo_iters:
                if it > 0:
                    self.say("EM (MAF) converged at iteration: " + str(int(it)))
            if self.rank0:
                suffix = "_downsampled" if self.L_ds is not None else ""
                outfile = "%s.pop_like_LOO%s.tsv" % (a.out, suffix)
                utils.write_ass_mats(outfile, ll, self.sample_names, pops, print_part_column=False,
                                     sample_locations=IDs[:, 1], doing_LOO=True)
                print("Saved leave-one-out cross validation log likelihoods as %s" % outfile)
                if a.partition_sites > 1:
                    partfile = "%s.pop_like_LOO%s_partitions_%d.tsv.gz" % (a.out, suffix, a.partition_sites)
                    utils.write_ass_mats(partfile, ll_parts, self.sample_names, pops, partition_count=a.partition_sites,
                                         print_part_column=True, sample_locations=IDs[:, 1], doing_LOO=True)
                    print("Saved leave-one-out cross validation log likelihoods from partitioned sites as %s" % partfile)
                print("Column order of populations is: %s" % pops)

    # -- --get_pop_like ----------------------------------------------------------------------
    def pop_like(self):
        np, a = self.np, self.args
        from . import glassy
        self.say("Parsing population allele frequency file.")
        assert os.path.isfile(a.pop_af_file), "Population allele frequency file does not exist!!"
        A = self.shard(np.load(a.pop_af_file))
        self.say("Calculating likelihood of population assignment")
        ll = glassy.assignLL(self.L, A, a.threads)
        if self.rank0:
            np.savetxt(a.out + ".pop_like.txt", ll, fmt="%.7f")
            print("Saved population assignment log likelihoods as " + str(a.out) + ".pop_like.txt (text)")

    # -- z-scores ------------------------------------------------------------------------------
    def zscores(self, reference_mode):
        np, a = self.np, self.args
        from . import zscore
        self.say("Parsing population ID file.")
        assert os.path.isfile(a.pop_af_IDs), "Population ID file does not exist!!"
        IDs = np.loadtxt(a.pop_af_IDs, delimiter="\t", dtype="str")
        A = None
        if not reference_mode:
            self.say("Parsing population allele frequency file.")
            assert os.path.isfile(a.pop_af_file), "Population allele frequency file does not exist!!"
            A = self.shard(np.load(a.pop_af_file))
        self.say("Parsing individual allele depths file.")
        assert os.path.isfile(a.ind_ad_file), "Individual allele depths file does not exist!"
        from . import reader
        rows = None
        if self._world() > 1:
            rows = self._range(self.M_total if self.M_total is not None else self.L.shape[0])
        # text (plain or gzipped) through the streaming parser, straight to saturating uint8 pairs; a rank reads only its rows
        AD = reader.readAD(a.ind_ad_file, a.threads, rows=rows)
        if AD.dtype != np.uint8:
            AD = np.ascontiguousarray(AD, dtype=np.int32)
        deep = int(np.count_nonzero(AD >= 255)) if AD.dtype == np.uint8 else int(np.count_nonzero(AD > 254))
        if deep and self.rank0:
            print("Note: " + str(deep) + " allele depths of 255 or more: those sites are treated as deeper than any depth class "
                  "(the reference never keeps them either)")
        assert os.path.isfile(a.pop_names), "Population names file does not exist!!"
        pops = np.loadtxt(a.pop_names, dtype="str")
        n = self.L.shape[1] // 2
        assert (n == IDs.shape[0]), "Number of individuals in beagle and reference ID file do not match!"
        thr = 0
        if a.allele_count_threshold is not None:
            assert (a.allele_count_threshold >= 0), "Allele count threshold needs to be greater than/equal to 0!"
            thr = a.allele_count_threshold
        start, end = 0, n
        if a.ind_start is not None:        # an explicit 0 is rejected, like the reference (WGSassign.py:335)
            assert (a.ind_start > 0 and a.ind_start <= n), "Start individual index needs to be within range of number of individuals!"
            start = a.ind_start
        if a.ind_end is not None:
            assert (a.ind_end > 0 and a.ind_end <= n), "End individual index needs to be within range of number of individuals!"
            end = a.ind_end
        rows = zscore.zscore_all(self.L, AD, IDs, zscore.MODE_REFERENCE if reference_mode else zscore.MODE_ASSIGNMENT,
                                 A=A, pops=pops, n_threshold=thr, single_read=a.single_read_threshold,
                                 ind_start=start, ind_end=end, maf_iter=a.maf_iter, maf_tole=a.maf_tole)
        z_out = np.empty((end - start, 1), dtype=np.float32)
        for j, r in enumerate(rows):
            if reference_mode and r["em_iters"] > 0:
                self.say("EM (MAF) converged at iteration: " + str(r["em_iters"]))
            self.say("Finished individual " + str(start + j))
            self.say("z_mu: " + str(r["z_mu"]))
            self.say("z_var: " + str(r["z_var"]))
            self.say("z_obs: " + str(r["w_obs"]))
            self.say("Loci used: " + str(r["loci_kept"]))
            self.say("Z-score: " + str(r["z"]))
            z_out[j, 0] = r["z"]
        if self.rank0:
            name = ".reference_z_ind.txt" if reference_mode else ".z_ind.txt"
            np.savetxt(a.out + name, z_out, fmt="%.7f")
            print("Saved " + str(end - start) + " individual z-scores as " + str(a.out) + name + " (text)")

    # -- mixtures ------------------------------------------------------------------------------
    def mixtures(self, mcmc):
        np, a = self.np, self.args
        from . import mixture
        print("Parsing population assignment likelihood file.")
        assert os.path.isfile(a.pop_like), "Population assignment log likelihood file does not exist!!"
        assert os.path.isfile(a.pop_like_IDs), "ID file does not exist!!"
        ll = np.loadtxt(a.pop_like)
        index = np.loadtxt(a.pop_like_IDs, delimiter="\t", dtype="str")
        print("Calculating mixture proportions with EM")
        res = mixture.mcmc_mix(ll, index, a.mixture_iter) if mcmc else mixture.em_mix(ll, index, a.mixture_iter, logsumexp=a.em_mix_logsumexp)
        np.savetxt(a.out + ".em_mix.txt", res, fmt="%s")       # the reference writes .em_mix.txt in both modes (WGSassign.py:470)
        if mcmc:
            print("Saved MCMC mixture proportions " + str(a.out) + ".mcmc_mix.txt (text)")
        else:
            print("Saved EM mixture proportions " + str(a.out) + ".em_mix.txt (text)")


def main(argv=None):
    args = parser.parse_args(argv)
    if argv is None and len(sys.argv) < 2:
        parser.print_help()
        sys.exit()
    run = None
    if args.loo_downsampled_beagle and not args.loo:
        raise ValueError("The --loo_downsampled_beagle option requires that --loo is also specified.")
    run = _Run(args)
    run.say("WGSassign")
    run.say("Matt DeSaix.")
    run.say("Using " + str(args.threads) + " thread(s).\n")
    if run.rank0:
        _write_args_file(args)
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(args.threads)
    run.parse_inputs()
    if args.get_reference_af:
        run.reference_af()
    if args.get_pop_like:
        run.pop_like()
    if args.get_reference_z_score:
        run.zscores(reference_mode=True)
    if args.get_assignment_z_score:
        run.zscores(reference_mode=False)
    if run.rank0 and args.get_em_mix:
        run.mixtures(mcmc=False)
    if run.rank0 and args.get_mcmc_mix:
        run.mixtures(mcmc=True)


if __name__ == "__main__":
    main()
