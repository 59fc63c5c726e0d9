"""Host-side helpers with the reference's surface (utils.py:9-151): name preview, site
intersection for the down-sampled LOO, the LOO TSV writer and modulo partition sums."""
import gzip

import numpy as np


def print_sample_and_site_summary(sample_names, site_names):
    """Same two lines as the reference prints after parsing (utils.py:9-18)."""
    def short(names):
        names = list(names)
        if len(names) <= 4:
            return ", ".join(names)
        return ", ".join(names[:2]) + ", ..., " + ", ".join(names[-2:])
    print("sample_names: %d samples total: %s" % (len(sample_names), short(sample_names)))
    print("site_names: %d sites total: %s" % (len(site_names), short(site_names)))


def filter_sites_to_common(L, site_names, site_names_target):
    """Rows of ``L`` (and their names) whose site is in ``site_names_target`` (utils.py:22-42)."""
    names = np.asarray(site_names)
    keep = np.isin(names, list(set(site_names_target)))
    dropped = int(np.count_nonzero(~keep))
    if dropped > 0:
        print("\tFiltered out %d sites not present in the target site list." % dropped)
    return L[keep, :], names[keep].tolist()


def write_ass_mats(filename, loglike_mat, sample_names, pop_names, partition_count=1, print_part_column=True,
                   sample_locations=None, doing_LOO=False):
    """Tab-separated assignment matrix, ``%.6f``, gzip when the name ends in .gz
    (utils.py:49-123: columns sample[, source_pop|location][, data_part], then populations)."""
    import pandas as pd
    n_ind, K = len(sample_names), len(pop_names)
    if loglike_mat.shape != (n_ind * partition_count, K):
        raise ValueError("loglike_mat shape mismatch: expected %s, got %s" % ((n_ind * partition_count, K), loglike_mat.shape))
    if not print_part_column and partition_count != 1:
        raise ValueError("print_part_column=False is only allowed if partition_count == 1")
    if sample_locations is not None:
        if len(sample_locations) != n_ind:
            raise ValueError("Length of sample_locations does not match sample_names")
        if doing_LOO and not set(sample_locations).issubset(set(pop_names)):
            raise ValueError("sample_locations contains values not in pop_names (required for LOO mode)")
    cols = {"sample": np.repeat(sample_names, partition_count)}
    if sample_locations is not None:
        cols["source_pop" if doing_LOO else "location"] = np.repeat(sample_locations, partition_count)
    if print_part_column:
        cols["data_part"] = np.tile(np.arange(partition_count), n_ind)
    table = pd.concat([pd.DataFrame(cols), pd.DataFrame(loglike_mat, columns=pop_names)], axis=1)
    if filename.endswith(".gz"):
        with gzip.open(filename, "wt") as fh:
            table.to_csv(fh, sep="\t", index=False, float_format="%.6f")
    else:
        table.to_csv(filename, sep="\t", index=False, float_format="%.6f")
    print("Wrote assignment matrix to %s" % filename)


def partition_loglikes(per_site_ll, partition_count):
    """Per-partition sums of a per-site vector, partition = site index mod count
    (utils.py:129-151).  Host helper; the LOO path computes these sums on the GPU."""
    per_site_ll = np.asarray(per_site_ll)
    if per_site_ll.ndim != 1:
        raise ValueError("per_site_ll must be a 1D array")
    out = np.zeros(partition_count, np.float32)
    np.add.at(out, np.arange(per_site_ll.shape[0]) % partition_count, per_site_ll)
    return out
