// oracle/oracle_msa.cpp -- TEST INFRASTRUCTURE ONLY (CPU restatement; never linked into the product).
//
// Restatement of the reference's progressive sum-of-pairs aligner
//   MultipleSequenceAlignmentSP<Index2D,SimpleScoreModel,vector,string,char>
//     ::align    /root/reference/StrainCall/MultipleSequenceAlignmentSP.cpp:10-49
//     ::forward  MultipleSequenceAlignmentSP.cpp:52-249
//     ::backward MultipleSequenceAlignmentSP.cpp:252-301
// scored by SimpleDnaScore (SimpleDnaScore.cpp:16-42, defaults Score.hpp:33).
// Parity status: PINNED against oracle/_ref (the reference sources compiled unmodified) by
// tests/test_oracle_vs_ref.py and against the committed vectors in tests/golden/msa_*.json.
//
// Deliberately kept close to the reference's arithmetic (per-sequence loops, one state per
// sequence and cell) so that it restates the reference, quirks included:
//  * `s` (sequences already in the profile) is read uninitialised by the reference (line 13,24);
//    the value its algorithm needs -- and that the pinned build uses -- is 1 before the 2nd input.
//  * when a cell is taken as match or delete, the per-sequence state PP is derived from the FIRST
//    row of the profile column for every sequence (the iterator it3 is never advanced,
//    lines 208-218 and 235-245); the first column (j==0) does it per row (lines 131-136).
//  * unknown letters score 0 against everything (std::map::operator[] default, SimpleDnaScore.cpp:11-14).
#include "oracle.h"

namespace oracle {

static bool known(char c)
{
    switch (c)
    {
        case 'A': case 'a': case 'C': case 'c': case 'G': case 'g': case 'T': case 't': case '+': case '-':
            return true;
    }
    return false;
}

// SimpleDnaScore::set with match=3, mismatch=-5, gap_open=-4, gap_extend=-2
int sp_score(char x, char y)
{
    if (!known(x) || !known(y)) return 0;
    if (x == y) return 3;
    if ((x ^ y) == 0x20 && (x | 0x20) >= 'a' && (x | 0x20) <= 't') return 3;  // same base, other case
    if ((x == '+' && y == '-') || (x == '-' && y == '+')) return 3;
    if (x == '+' || y == '+') return -4 + -2;
    if (x == '-' || y == '-') return -2;
    return -5;
}

namespace {
enum { MAT = 0, INS = 1, DEL = 2 };

// profile = vector of columns, each column holds s letters
void forward_backward(const std::string& seq, std::vector<std::string>& prof, int s)
{
    const int m = (int)prof.size() + 1;
    const int n = (int)seq.size() + 1;
    std::vector<long long> SC((size_t)m * n);
    std::vector<int> SI((size_t)m * n), SJ((size_t)m * n);
    std::vector<unsigned char> PP((size_t)m * n * s);
    auto pp = [&](int i, int j, int k) -> unsigned char& { return PP[((size_t)i * n + j) * s + k]; };

    SC[0] = 0; SI[0] = 0; SJ[0] = 0;
    for (int k = 0; k < s; ++k) pp(0, 0, k) = MAT;
    for (int j = 1; j < n; ++j)
    {
        long long sp = 0;
        for (int k = 0; k < s; ++k) sp += sp_score('A', j == 1 ? '+' : '-');
        SC[j] = SC[j - 1] + sp; SI[j] = 0; SJ[j] = -1;
        for (int k = 0; k < s; ++k) pp(0, j, k) = INS;
    }
    for (int i = 1; i < m; ++i)
    {
        const std::string& col = prof[i - 1];
        long long sp = 0;
        for (int k = 0; k < s; ++k) sp += sp_score(col[k], i == 1 ? '+' : '-');
        SC[(size_t)i * n] = SC[(size_t)(i - 1) * n] + sp; SI[(size_t)i * n] = -1; SJ[(size_t)i * n] = 0;
        for (int k = 0; k < s; ++k) pp(i, 0, k) = (col[k] == '-') ? MAT : DEL;
    }
    for (int i = 1; i < m; ++i)
    {
        const std::string& col = prof[i - 1];
        for (int j = 1; j < n; ++j)
        {
            const char c = seq[j - 1];
            long long r1 = 0, r2 = 0, r3 = 0;
            for (int k = 0; k < s; ++k)
            {
                if (col[k] == '-') r1 += sp_score(pp(i - 1, j - 1, k) == INS ? '-' : '+', c);
                else r1 += sp_score(col[k], c);
            }
            r1 += SC[(size_t)(i - 1) * n + (j - 1)];
            for (int k = 0; k < s; ++k) r2 += sp_score(pp(i, j - 1, k) == INS ? '-' : '+', c);
            r2 += SC[(size_t)i * n + (j - 1)];
            for (int k = 0; k < s; ++k)
            {
                if (col[k] != '-') r3 += sp_score(col[k], pp(i - 1, j, k) == DEL ? '-' : '+');
                else r3 += sp_score(col[k], '-');
            }
            r3 += SC[(size_t)(i - 1) * n + j];

            const size_t at = (size_t)i * n + j;
            if (r1 >= r2 && r1 >= r3)
            {
                SC[at] = r1; SI[at] = -1; SJ[at] = -1;
                for (int k = 0; k < s; ++k) pp(i, j, k) = (col[0] == '-') ? INS : MAT;
            }
            else if (r2 >= r1 && r2 >= r3)
            {
                SC[at] = r2; SI[at] = 0; SJ[at] = -1;
                for (int k = 0; k < s; ++k) pp(i, j, k) = INS;
            }
            else
            {
                SC[at] = r3; SI[at] = -1; SJ[at] = 0;
                for (int k = 0; k < s; ++k) pp(i, j, k) = (col[0] == '-') ? MAT : DEL;
            }
        }
    }
    // traceback (backward): rebuild the profile with one more row
    std::vector<std::string> rev;
    int x = m - 1, y = n - 1;
    int pi = (int)prof.size() - 1, sj = (int)seq.size() - 1;
    while (!(x == 0 && y == 0))
    {
        const int di = SI[(size_t)x * n + y], dj = SJ[(size_t)x * n + y];
        if (di == -1 && dj == -1) { std::string t = prof[pi--]; t.push_back(seq[sj--]); rev.push_back(t); }
        else if (di == 0 && dj == -1) { std::string t(s, '-'); t.push_back(seq[sj--]); rev.push_back(t); }
        else { std::string t = prof[pi--]; t.push_back('-'); rev.push_back(t); }
        x += di; y += dj;
    }
    prof.assign(rev.rbegin(), rev.rend());
}
}  // namespace

// rows[t] = the canonised form of seqs[t]; every row has the same length (the profile width)
std::vector<std::string> msa_sp_align(const std::vector<std::string>& seqs)
{
    std::vector<std::string> prof;
    if (seqs.empty()) return {};
    for (char c : seqs[0]) prof.push_back(std::string(1, c));
    int s = 1;
    for (size_t t = 1; t < seqs.size(); ++t, ++s) forward_backward(seqs[t], prof, s);
    std::vector<std::string> rows(seqs.size());
    for (size_t t = 0; t < seqs.size(); ++t)
        for (const std::string& col : prof)
            if (t < col.size()) rows[t].push_back(col[t]);
    return rows;
}

}  // namespace oracle
