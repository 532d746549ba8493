"""Reference 'N' runs for LeadProvider._mask_N_coverage (leadprov.py:420-443, only with --reference):
a plain FASTA scan (no pysam), producing the per-task intervals `RecordBlock.set_n_mask` expects."""
import gzip


def n_runs(fasta_path, wanted=None):
    """{contig name: [(start, end), ...]} of maximal runs of 'N' (upper case only, as the reference compares == 78)."""
    opener = gzip.open if str(fasta_path).endswith(".gz") else open
    out, name, pos, run = {}, None, 0, None
    with opener(fasta_path, "rt") as f:
        for line in f:
            if line.startswith(">"):
                if name is not None and run is not None:
                    out[name].append((run, pos))
                name, pos, run = line[1:].split()[0], 0, None
                if wanted is None or name in wanted:
                    out[name] = []
                else:
                    name = None
                continue
            if name is None:
                continue
            s = line.rstrip("\n")
            i = 0
            while i < len(s):
                if s[i] == "N":
                    if run is None:
                        run = pos + i
                    j = i
                    while j < len(s) and s[j] == "N":
                        j += 1
                    i = j
                    if j < len(s):
                        out[name].append((run, pos + j))
                        run = None
                else:
                    k = s.find("N", i)
                    i = len(s) if k < 0 else k
            pos += len(s)
    if name is not None and run is not None:
        out[name].append((run, pos))
    return out
