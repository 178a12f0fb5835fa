"""Restatement of cutadapt's semiglobal aligner (upstream src/cutadapt/_align.pyx,
``Aligner.__cinit__`` / ``Aligner.locate``; src/cutadapt/align.py ``EndSkip``).

Unit-cost edit distance decides whether a match is allowed; a score (+1 match,
-1 mismatch, -2 insertion/deletion) decides between candidates; ``origin`` tracks where
in the read the path started (negative: that many adapter characters were skipped).
One column of (cost, score, origin) is kept, with Ukkonen's band, exactly as upstream.
Only the plain-ASCII comparison is restated: the reference only ever builds adapters
from [ACGT]+ (common.py:174, run.py:1056) for which cutadapt switches wildcards off.
"""

from enum import IntFlag

MATCH_SCORE = +1
MISMATCH_SCORE = -1
INSERTION_SCORE = -2
DELETION_SCORE = -2


class EndSkip(IntFlag):
    REFERENCE_START = 1  # a prefix of the reference may be skipped at no cost
    QUERY_START = 2      # a prefix of the query may be skipped at no cost
    REFERENCE_END = 4    # a suffix of the reference may be skipped at no cost
    QUERY_STOP = 8       # a suffix of the query may be skipped at no cost
    SEMIGLOBAL = 15


class Aligner:
    def __init__(self, reference, max_error_rate, flags=15, wildcard_ref=False,
                 wildcard_query=False, indel_cost=1, min_overlap=1):
        if wildcard_ref or wildcard_query:
            raise NotImplementedError("wildcard matching is never active on the cutseq path")
        if indel_cost != 1:
            raise NotImplementedError("cutseq always aligns with indels=True (cost 1)")
        if min_overlap < 1:
            raise ValueError("minimum overlap must be at least 1")
        self.reference = reference
        self.m = len(reference)
        if self.m == 0:
            raise ValueError("reference must not be empty")
        self.max_error_rate = max_error_rate
        self.start_in_reference = bool(flags & 1)
        self.start_in_query = bool(flags & 2)
        self.stop_in_reference = bool(flags & 4)
        self.stop_in_query = bool(flags & 8)
        self.min_overlap = min_overlap

    def locate(self, query):
        """-> (refstart, refstop, querystart, querystop, score, errors) or None"""
        s1 = self.reference
        s2 = query
        m = self.m
        n = len(query)
        max_error_rate = self.max_error_rate
        k = int(max_error_rate * m)

        max_n = n
        min_n = 0
        if not self.start_in_query:
            max_n = min(n, m + k)
        if not self.stop_in_query:
            min_n = max(0, n - m - k)

        cost = [0] * (m + 1)
        score = [0] * (m + 1)
        origin = [0] * (m + 1)
        if not self.start_in_reference and not self.start_in_query:
            for i in range(m + 1):
                cost[i] = max(i, min_n)
                origin[i] = 0
        elif self.start_in_reference and not self.start_in_query:
            for i in range(m + 1):
                cost[i] = min_n
                origin[i] = min(0, min_n - i)
        elif not self.start_in_reference and self.start_in_query:
            for i in range(m + 1):
                cost[i] = i
                origin[i] = max(0, min_n - i)
        else:
            for i in range(m + 1):
                cost[i] = min(i, min_n)
                origin[i] = min_n - i

        NONE = m + n + 1
        best_cost, best_origin, best_score = NONE, 0, 0
        best_ref_stop, best_query_stop = m, n

        last = min(m, k + 1)
        if self.start_in_reference:
            last = m

        for j in range(min_n + 1, max_n + 1):
            d_cost, d_score, d_origin = cost[0], score[0], origin[0]
            if self.start_in_query:
                origin[0] = j
            else:
                cost[0] = j
            c2 = s2[j - 1]
            for i in range(1, last + 1):
                if s1[i - 1] == c2:
                    c, o, s = d_cost, d_origin, d_score + MATCH_SCORE
                else:
                    cost_diag = d_cost + 1
                    cost_deletion = cost[i] + 1
                    cost_insertion = cost[i - 1] + 1
                    if cost_diag <= cost_deletion and cost_diag <= cost_insertion:
                        c, o, s = cost_diag, d_origin, d_score + MISMATCH_SCORE
                    elif cost_insertion <= cost_deletion:
                        c, o, s = cost_insertion, origin[i - 1], score[i - 1] + INSERTION_SCORE
                    else:
                        c, o, s = cost_deletion, origin[i], score[i] + DELETION_SCORE
                d_cost, d_score, d_origin = cost[i], score[i], origin[i]
                cost[i], score[i], origin[i] = c, s, o
            while last >= 0 and cost[last] > k:
                last -= 1
            if last < m:
                last += 1
            elif self.stop_in_query:
                c, s, o = cost[m], score[m], origin[m]
                length = m + min(o, 0)
                is_acceptable = length >= self.min_overlap and c <= length * max_error_rate
                best_length = m + min(best_origin, 0)
                if is_acceptable and (
                    best_cost == NONE
                    or (o <= best_origin + m // 2 and s > best_score)
                    or (length > best_length and s > best_score)
                ):
                    best_score, best_cost, best_origin = s, c, o
                    best_ref_stop, best_query_stop = m, j
                    if c == 0 and o >= 0:
                        break

        if max_n == n:
            first_i = 0 if self.stop_in_reference else m
            for i in range(m, first_i - 1, -1):
                length = i + min(origin[i], 0)
                c, s = cost[i], score[i]
                is_acceptable = length >= self.min_overlap and c <= length * max_error_rate
                if is_acceptable and (s > best_score or (s == best_score and c < best_cost)):
                    best_score, best_cost, best_origin = s, c, origin[i]
                    best_ref_stop, best_query_stop = i, n

        if best_cost == NONE:
            return None
        if best_origin >= 0:
            start1, start2 = 0, best_origin
        else:
            start1, start2 = -best_origin, 0
        return (start1, best_ref_stop, start2, best_query_stop, best_score, best_cost)
