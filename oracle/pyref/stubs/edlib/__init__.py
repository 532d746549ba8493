"""Stand-in for the `edlib` package (absent from this container) so that the UNMODIFIED reference runs its SVGroup.align_call path
(sv.py:24,282-292).  The reference only reads align(a, b)['editDistance'] with edlib's defaults (mode "NW" = global alignment, task "distance",
no threshold): the Levenshtein distance, a uniquely defined number, computed here with Myers' bit-vector recurrence on Python integers
(checked against the textbook dynamic programme in tests/test_combine.py)."""


def levenshtein(a, b):
    if len(a) > len(b):
        a, b = b, a
    m = len(a)
    if m == 0:
        return len(b)
    peq = {}
    for i, ch in enumerate(a):
        peq[ch] = peq.get(ch, 0) | (1 << i)
    mask, top = (1 << m) - 1, 1 << (m - 1)
    pv, mv, score = mask, 0, m
    for ch in b:
        eq = peq.get(ch, 0)
        xv = eq | mv
        xh = (((eq & pv) + pv) ^ pv) | eq
        ph = (mv | ~(xh | pv)) & mask
        mh = pv & xh
        if ph & top:
            score += 1
        elif mh & top:
            score -= 1
        ph = ((ph << 1) | 1) & mask
        mh = (mh << 1) & mask
        pv = (mh | ~(xv | ph)) & mask
        mv = ph & xv
    return score


def align(query, target, mode="NW", task="distance", k=-1, additionalEqualities=None):
    if mode != "NW" or k != -1 or additionalEqualities:
        raise NotImplementedError("the stand-in covers the defaults the reference uses")
    return {"editDistance": levenshtein(query, target), "alphabetLength": len(set(query) | set(target)), "locations": [(None, len(target) - 1)], "cigar": None}
