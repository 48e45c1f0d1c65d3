"""Test helper: minimal reader for the reference's raw matrix format (.cn) and writer for its
segment table (.seg), restating lib/RawSampleSet.hpp:217-285 (+ sort :332-386), lib/global.hpp:62-90
(chromosome names) and lib/SegmentedSampleSet.hpp:519-535.  Used to drive the oracle the way
`cna segment` drives lib/cbs; the product has its own C++ implementation (genomic_b200/host)."""
from __future__ import annotations

import numpy as np

N_CHROM = 24


def chrom_index(name: str) -> int:
    """1..24, 0 for unknown (rows with unknown chromosome are skipped by the reference)."""
    s = name[3:] if name.startswith("chr") else name
    if s == "X":
        return 23
    if s == "Y":
        return 24
    if s.isdigit() and not (len(s) > 1 and s[0] == "0"):
        v = int(s)
        return v if 1 <= v <= N_CHROM else 0
    return 0


def _parse_float32(tok: str):
    """std::from_chars<float> accepts no leading '+' or whitespace; 'NA' etc. fail and are skipped."""
    if not tok or tok[0] == "+" or tok != tok.strip():
        return None
    low = tok.lower().lstrip("-")
    if low in ("nan", "inf", "infinity"):
        return np.float32(tok)
    try:
        return np.float32(tok)
    except ValueError:
        return None


def read_cn(path: str):
    names = []
    pos = [[] for _ in range(N_CHROM)]
    vals = None
    with open(path) as f:
        lines = f.read().split("\n")
    if lines and lines[-1] == "":
        lines.pop()  # getline + eof check: a final unterminated line is dropped only if empty
    for ln, line in enumerate(lines):
        fields = line.split("\t")
        if ln == 0:
            names = fields[3:]
            vals = [[[] for _ in range(N_CHROM)] for _ in names]
            continue
        if len(fields) < 3:
            continue
        try:
            p = int(fields[2])
        except ValueError:
            continue
        c = chrom_index(fields[1])
        if c == 0:
            continue
        pos[c - 1].append(p)
        k = 0
        for tok in fields[3:]:
            v = _parse_float32(tok)
            if v is None:
                continue  # skipped fields shift later samples' columns, as in the reference
            if k < len(names):
                vals[k][c - 1].append(v)
            k += 1
    # per chromosome sort by position (stable order of equal positions is not pinned by the reference)
    out_pos, out_vals = [], [[None] * N_CHROM for _ in names]
    for c in range(N_CHROM):
        p = np.asarray(pos[c], dtype=np.int64)
        order = np.argsort(p, kind="stable")
        out_pos.append(p[order])
        for s in range(len(names)):
            v = np.asarray(vals[s][c], dtype=np.float32)
            out_vals[s][c] = v[order] if len(v) == len(order) else v
    return names, out_pos, out_vals


def is_log_scale(vals) -> bool:
    """src/cna_segment.hpp:109-125."""
    neg = pos = False
    for sample in vals:
        for v in sample:
            v = np.asarray(v, dtype=np.float64)
            v = v[np.isfinite(v)]
            neg |= bool((v < 0).any())
            pos |= bool((v > 0).any())
    return neg and pos


def fmt_float(v) -> str:
    """default std::ostream formatting of a float: %g with 6 significant digits."""
    return "%g" % float(np.float32(v))


def seg_text(names, positions, units, seg_count, lengths, means) -> str:
    """units: list of (sample_index, chrom_index0) in processing order."""
    rows = ["sample\tchromosome\tstart\tend\tcount\tstate"]
    k = 0
    for u, (s, c) in enumerate(units):
        start = 0
        for _ in range(int(seg_count[u])):
            ln = int(lengths[k])
            if ln > 0:
                end = start + ln - 1
                rows.append("%s\t%d\t%d\t%d\t%d\t%s" % (names[s], c + 1, positions[c][start], positions[c][end], ln,
                                                       fmt_float(means[k])))
                start += ln
            k += 1
    return "\n".join(rows) + "\n"


def cohort_from_cn(path: str):
    """Flatten a .cn file into (values float64, unit_off, chrom_label, units) in `cna segment` order:
    samples in file order, chromosomes 1..24, empty chromosomes skipped."""
    names, positions, vals = read_cn(path)
    chunks, off, lab, units = [], [0], [], []
    for s in range(len(names)):
        for c in range(N_CHROM):
            v = vals[s][c]
            if len(v) == 0:
                continue
            chunks.append(np.asarray(v, dtype=np.float32).astype(np.float64))
            off.append(off[-1] + len(v))
            lab.append(c + 1)
            units.append((s, c))
    values = np.concatenate(chunks) if chunks else np.zeros(0)
    return names, positions, vals, values, np.asarray(off, dtype=np.int64), np.asarray(lab, dtype=np.int32), units
