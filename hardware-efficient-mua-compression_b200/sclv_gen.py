"""Candidate-SCLV generator: regenerates the `Stored_SCLVs_S_<S>.pkl` tables of the
reference's `Compressing data/Produce SCLVs/produce_all_SCLVs_given_S.py` -- rows AND row order -- and, unlike the
reference, also keeps the codeword strings of the first Huffman code that produced each row (the reference reduces
them to lengths at :77,:87-89).

Offline, one-shot host tool like the reference's (the GPU path consumes the tables as data).  What has to be
reproduced exactly for the row order to match (:39-98):
  * the probability grid walk: find every position equal to the current minimum of p, zero all but the first of them
    and add 0.15 to that first one, until the last position reaches 1 (:55-68).  Grid values are repeated float64
    additions of 0.15, so equal values are bit-identical and `==` is exact;
  * Huffman with a binary heap whose items order by (weight, leaf list) -- ties on the weight fall through to the
    leaves' [symbol, code] lists exactly as Python compares the reference's nested lists (:18-29);
  * codes sorted by (length, [symbol, code]) (:29); a row is kept the first time its length vector appears (:91-98).
"""
import heapq


def huffman_sorted(weights):
    """[(symbol, code)] sorted by (len(code), symbol, code) for float weights indexed by symbol."""
    heap = [(w, [[sym, ""]]) for sym, w in enumerate(weights)]
    heapq.heapify(heap)
    while len(heap) > 1:
        w_lo, lo = heapq.heappop(heap)
        w_hi, hi = heapq.heappop(heap)
        for leaf in lo:
            leaf[1] = "0" + leaf[1]
        for leaf in hi:
            leaf[1] = "1" + leaf[1]
        heapq.heappush(heap, (w_lo + w_hi, lo + hi))
    leaves = heap[0][1]
    return sorted(((s, c) for s, c in leaves), key=lambda sc: (len(sc[1]), [sc[0], sc[1]]))


def generate(S, increment=0.15):
    """-> (rows, codebooks): rows = list of ascending length vectors in first-seen order (the pickle's content),
    codebooks[i] = codeword strings (rank order) of the first Huffman code that gave rows[i]."""
    p = [0.0] * S
    rows, books, seen = [], [], set()
    while p[-1] < 1:
        lo = min(p)
        first = None
        for i, v in enumerate(p):
            if v == lo:
                if first is None:
                    first = i
                else:
                    p[i] = 0.0
        p[first] += increment
        total = sum_like_numpy(p)
        code = huffman_sorted([v / total for v in p])
        lens = tuple(len(c) for _, c in code)
        if lens not in seen:
            seen.add(lens)
            rows.append(list(lens))
            books.append([c for _, c in code])
    return rows, books


def sum_like_numpy(p):
    """np.sum of a short float64 vector: plain left-to-right accumulation for n < 8 elements, NumPy's
    8-accumulator unrolled pairwise kernel above that (n <= 10 here: r[0..7] then the tail added in order)."""
    n = len(p)
    if n < 8:
        t = 0.0
        for v in p:
            t += v
        return t
    r = list(p[:8])
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    for v in p[8:]:
        res += v
    return res


def generate_all(S_values=range(2, 11)):
    return {S: generate(S) for S in S_values}
