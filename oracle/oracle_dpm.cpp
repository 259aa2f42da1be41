// oracle/oracle_dpm.cpp -- TEST INFRASTRUCTURE ONLY (CPU restatement; never linked into the product).
//
// Restatement of the reference's Dirichlet-process strain clustering over the graph levels,
//   Strain model                   /root/reference/StrainCall/Strain.cpp:9-251
//   hard_clustering                NonparametricClustering.cpp:17-125
//   np_bayes_clustering            NonparametricClustering.cpp:127-244
//   streaming_clustering           NonparametricClustering.cpp:262-582
//   merge_strains / read_reassign  NonparametricClustering.cpp:584-702
//   read_assign (AlignRead form)   NonparametricClustering.cpp:776-836
// in the reference's own precision (x87 long double) and with the reference's random stream:
// std::mt19937(1234) re-created per call, std::discrete_distribution<> as libstdc++ 13 draws it
// (bits/random.tcc: weights converted to double, normalised by their sum, partial sums with the
// last forced to 1.0, one generate_canonical<double,53> = (w0 + w1*2^32)/2^64 per draw, result =
// lower_bound of the draw; NO draw at all when there are fewer than two weights).
// Parity status: PINNED against oracle/_ref by tests/test_oracle_vs_ref.py (paths, abundances,
// substitution tables and per-read log-likelihoods, all to the last printed digit).
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <limits>
#include <queue>
#include <random>
#include <set>

namespace oracle {

namespace {

const char kLetters[6] = {'A', 'C', 'G', 'T', '-', '='};
int letter_index(char c)
{
    for (int i = 0; i < 6; ++i) if (kLetters[i] == c) return i;
    return -1;
}
typedef std::pair<std::string, std::string> SubKey;
typedef std::map<SubKey, LD> SubCounts;

struct Strain
{
    LD Z = 0, abundance = 0;
    LD sub[6][6];
    LD comp[6];
    SubCounts other;  // keys outside the 6x6 table (std::map::operator[] would create them)
    std::map<int, LD> ll;
    std::vector<int> path;

    Strain() {}
    Strain(int N, LD e)
    {
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) sub[i][j] = (i == j) ? N * (1 - e) : N * e;
        totals();
    }
    void totals()
    {
        Z = 0;
        for (int i = 0; i < 6; ++i)
        {
            comp[i] = 0;
            for (int j = 0; j < 6; ++j) comp[i] += sub[i][j];
            Z += comp[i];
        }
    }
    // Strain::logprob(string,string), Strain.cpp:130-133
    LD logprob(const std::string& a, const std::string& b)
    {
        int ia = a.size() == 1 ? letter_index(a[0]) : -1;
        int ib = b.size() == 1 ? letter_index(b[0]) : -1;
        LD num, den;
        if (ia >= 0 && ib >= 0) num = sub[ia][ib];
        else
        {
            auto it = other.find(SubKey(a, b));
            num = it == other.end() ? (LD)0 : it->second;
        }
        den = ia >= 0 ? comp[ia] : (LD)0;
        return std::log(num) - std::log(den);
    }
    LD& loglik(int id) { return ll[id]; }  // Strain::logprob(int): operator[] creates a 0 entry
    void add_loglik(int id, LD x)
    {
        auto it = ll.find(id);
        if (it == ll.end()) ll[id] = x; else it->second += x;
    }
    // Strain::update_model, Strain.cpp:108-123
    void update(LD al, const SubCounts& sc)
    {
        abundance += al;
        for (auto& kv : sc)
        {
            const std::string& a = kv.first.first;
            const std::string& b = kv.first.second;
            int ia = a.size() == 1 ? letter_index(a[0]) : -1;
            int ib = b.size() == 1 ? letter_index(b[0]) : -1;
            if (ia >= 0 && ib >= 0) sub[ia][ib] += kv.second; else other[kv.first] += kv.second;
        }
        totals();
    }
};

void normalize(std::vector<LD>& f)
{
    LD z = 0;
    for (LD x : f) z += x;
    for (LD& x : f) x /= z;
}

// std::discrete_distribution<>(p.begin(),p.end())(gen) as libstdc++ evaluates it
struct Sampler
{
    std::mt19937 gen;
    Sampler() : gen(1234) {}
    int draw(const std::vector<LD>& w)
    {
        if (w.size() < 2) return 0;
        std::vector<double> prob(w.begin(), w.end());
        double sum = 0.0;
        for (double x : prob) sum += x;
        for (double& x : prob) x /= sum;
        std::vector<double> cp(prob.size());
        double run = 0.0;
        for (size_t i = 0; i < prob.size(); ++i) { run = (i == 0) ? prob[0] : run + prob[i]; cp[i] = run; }
        cp.back() = 1.0;
        double lo = (double)gen(), hi = (double)gen();
        double u = (lo + hi * 4294967296.0) / 18446744073709551616.0;
        if (u >= 1.0) u = std::nextafter(1.0, 0.0);
        return (int)(std::lower_bound(cp.begin(), cp.end(), u) - cp.begin());
    }
};

struct LevelRead { int rid; std::string s; int cn; };

const std::string& last_label(const Pog& g, const Strain& s) { return g.store[s.path.back()].label; }

std::string seq_of(const Pog& g, const std::vector<int>& path)
{
    std::string r;
    for (int h : path) r += g.store[h].label;
    return r;
}

int mate_of(const PairTable& pt, int id, int copy /* cn-1 */)
{
    return pt.val[pt.off[id] + copy];
}

// hard_clustering, NonparametricClustering.cpp:17-125
void hard_clustering(const Pog& g, std::vector<Strain>& strains, const std::vector<LevelRead>& reads,
                     const std::vector<int>& new_reads, const PairTable& pairs)
{
    const size_t S = strains.size();
    std::vector<LD> abundance(S, 0);
    std::vector<SubCounts> subst(S);
    for (size_t ri = 0; ri < reads.size(); ++ri)
    {
        const LevelRead& r = reads[ri];
        for (int cn = r.cn; cn > 0; --cn)
        {
            int uid = mate_of(pairs, r.rid, cn - 1);
            std::vector<LD> p(S, 0);
            for (size_t si = 0; si < S; ++si) p[si] = strains[si].abundance;
            normalize(p);
            for (size_t si = 0; si < S; ++si)
            {
                p[si] = std::log(p[si]) + strains[si].loglik(r.rid);
                if (uid >= 0) p[si] += strains[si].loglik(uid);
                p[si] = std::exp(p[si]);
            }
            normalize(p);
            for (size_t si = 0; si < S; ++si)
            {
                abundance[si] += p[si];
                const std::string& sb = last_label(g, strains[si]);
                if (r.s.size() == 1) subst[si][SubKey(sb, r.s)] += p[si];
                else if (new_reads[ri])
                {
                    size_t i = sb.size(), j = r.s.size();
                    while (i > 0 && j > 0)
                    {
                        --i; --j;
                        subst[si][SubKey(std::string(1, sb[i]), std::string(1, r.s[j]))] += p[si];
                    }
                }
                else
                {
                    size_t i = 0, j = 0;
                    while (i < sb.size() && j < r.s.size())
                    {
                        subst[si][SubKey(std::string(1, sb[i]), std::string(1, r.s[j]))] += p[si];
                        ++i; ++j;
                    }
                }
            }
        }
    }
    for (size_t c = 0; c < S; ++c) strains[c].update(abundance[c], subst[c]);
}

// np_bayes_clustering, NonparametricClustering.cpp:127-244 (the k-posterior it also tallies is unused)
void np_bayes_clustering(const Pog& g, std::vector<Strain>& strains, const std::vector<LevelRead>& reads,
                         const PairTable& pairs, int n, std::vector<LD>& abundance)
{
    const int S = (int)strains.size(), m = (int)reads.size();
    std::vector<LD> a(S), p(S, 0);
    std::vector<SubCounts> sc(S);
    Sampler sampler;
    for (int s = 0; s < S; ++s) a[s] = strains[s].abundance;
    int read_size = 0;
    for (auto& r : reads) read_size += r.cn;
    n = std::min(n, 40000 / read_size);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j)
        {
            const int id = reads[j].rid;
            for (int cn = reads[j].cn; cn > 0; --cn)
            {
                for (int s = 0; s < S; ++s) p[s] = a[s];
                normalize(p);
                for (int s = 0; s < S; ++s)
                {
                    p[s] = std::log(p[s]) + strains[s].loglik(id);
                    int uid = mate_of(pairs, id, cn - 1);
                    if (uid >= 0 && strains[s].ll.count(uid) > 0) p[s] += strains[s].loglik(uid);
                    p[s] = std::exp(p[s]);
                }
                int c = sampler.draw(p);
                a[c] += 1;
                sc[c][SubKey(last_label(g, strains[c]), reads[j].s)] += 1;
            }
        }
    normalize(a);
    for (LD& x : a) x *= read_size;
    for (auto& m1 : sc) for (auto& kv : m1) kv.second /= n;
    abundance = a;
    for (int s = 0; s < S; ++s) strains[s].update(a[s], sc[s]);
}

// seq_identity, NonparametricClustering.cpp:584-612
LD seq_identity(const std::string& a, const std::string& b)
{
    int iden = 0, len = 0;
    for (size_t i = 0; i < a.size(); ++i)
    {
        char x = a[i], y = i < b.size() ? b[i] : '\0';
        if ((x == '-' || x == '=') && (y == '-' || y == '=')) continue;
        if (x == '^' && y == '^') continue;
        if (x == y) iden += 1;
        len += 1;
    }
    return (iden + 0.0) / len;
}

void sort_by_abundance(std::vector<Strain>& v)
{
    // same comparator on the same sequence as the reference's std::sort: same permutation, ties included
    std::vector<int> idx(v.size());
    for (size_t i = 0; i < idx.size(); ++i) idx[i] = (int)i;
    std::sort(idx.begin(), idx.end(), [&](int a, int b) { return v[a].abundance > v[b].abundance; });
    std::vector<Strain> r;
    for (int i : idx) r.push_back(v[i]);
    v.swap(r);
}

// merge_strains, NonparametricClustering.cpp:645-670
void merge_strains(const Pog& g, std::vector<Strain>& strains, LD diff)
{
    if (strains.empty()) return;
    sort_by_abundance(strains);
    std::vector<Strain> merged(1, strains[0]);
    for (size_t i = 1; i < strains.size(); ++i)
    {
        size_t j = 0;
        for (; j < merged.size(); ++j)
            if (seq_identity(seq_of(g, strains[i].path), seq_of(g, merged[j].path)) > 1 - diff)
            {
                merged[j].abundance += strains[i].abundance;
                break;
            }
        if (j == merged.size()) merged.push_back(strains[i]);
    }
    strains = merged;
}

}  // namespace

std::string strain_seq(const Pog& g, const StrainOut& s) { return seq_of(g, s.path); }
std::string plain_seq(const Pog& g, const StrainOut& s)
{
    std::string r;
    for (int h : s.path)
    {
        const std::string& l = g.store[h].label;
        if (l != "^" && l != "$" && l != "-" && l != "=") r += l;
    }
    return r;
}

// streaming_clustering, NonparametricClustering.cpp:262-582
void infer_strains(const Pog& g, const PairTable& pairs, int n, LD e, LD tau, LD diff, std::vector<StrainOut>& out)
{
    bool branching = false;
    std::vector<Strain> level_strains, next_strains, result;
    std::queue<int> cur, nxt;
    std::set<int> queued;
    std::vector<LevelRead> level_reads;
    int level_read_count = 0;
    std::set<std::pair<int, int>> total_reads;
    std::vector<int> new_reads;
    std::vector<LD> abundance;

    level_strains.push_back(Strain(100, e));
    cur.push(g.order[0]);
    while (!cur.empty())
    {
        const int u = cur.front(); cur.pop();
        const Node& nu = g.store[u];
        if (u == g.order[0])
        {
            level_strains[0].path.push_back(u);
            level_strains[0].abundance = 1;
        }
        else if (nu.label == "$")
        {
            // read_reassign: only its sort and its map look-ups leave a trace (lines 672-702)
            sort_by_abundance(level_strains);
            for (auto& rr : total_reads) for (auto& s : level_strains) s.loglik(rr.first);
            merge_strains(g, level_strains, diff);
            result = level_strains;
        }
        else
            for (auto& p : nu.pool) { level_reads.push_back({p.rid, p.s, p.cn}); level_read_count += p.cn; }

        for (int v : nu.out) if (!queued.count(v)) { nxt.push(v); queued.insert(v); }
        if (!cur.empty()) continue;

#ifdef ORACLE_TRACE
        {
            LD tot = 0; for (auto& s : level_strains) tot += s.abundance;
            fprintf(stderr, "level strains=%zu reads=%zu branching=%d absum=%Lg\n", level_strains.size(), level_reads.size(), (int)branching, tot);
        }
#endif
        // ---- end of a level: update the per-read log-likelihood of every candidate strain
        if (!level_reads.empty())
        {
            new_reads.assign(level_reads.size(), 0);
            for (auto& s : level_strains)
            {
                for (size_t ri = 0; ri < level_reads.size(); ++ri)
                {
                    const int rid = level_reads[ri].rid;
                    const std::string& rb = level_reads[ri].s;
                    std::string sb = last_label(g, s);
                    LD loglik;
                    if (sb.size() == 1)
                    {
                        if (sb == "N") sb = rb;
                        loglik = s.logprob(sb, rb);
                    }
                    else
                    {
                        loglik = 0;
                        if (s.ll.count(rid) == 0)
                        {
                            size_t ii = sb.size(), jj = rb.size();
                            while (ii > 0 && jj > 0)
                            {
                                std::string a(1, sb[--ii]), b(1, rb[--jj]);
                                if (a == "N") a = b;
                                loglik += s.logprob(a, b);
                            }
                            new_reads[ri] = 1;
                        }
                        else
                        {
                            size_t ii = 0, jj = 0;
                            while (ii < sb.size() && jj < rb.size())
                            {
                                std::string a(1, sb[ii++]), b(1, rb[jj++]);
                                if (a == "N") a = b;
                                loglik += s.logprob(a, b);
                            }
                        }
                    }
                    s.add_loglik(rid, loglik);
                }
            }
            for (auto& r : level_reads) total_reads.insert({r.rid, r.cn});
        }

        if (branching && !level_reads.empty())
        {
            std::map<std::string, LD> before, after;
            for (auto& s : level_strains) before[seq_of(g, s.path)] = s.abundance;
            np_bayes_clustering(g, level_strains, level_reads, pairs, n, abundance);
            for (auto& s : level_strains) after[seq_of(g, s.path)] = s.abundance;
            LD dmax = 0;
            for (auto& s : level_strains)
            {
                std::string q = seq_of(g, s.path);
                LD d = after[q] - before[q];
                if (dmax < d) dmax = d;
            }
            LD Z = 0;
            for (LD z : abundance) Z += z;
            const LD Zt = Z * tau;
            std::vector<size_t> drop;
            for (size_t si = 0; si < level_strains.size(); ++si)
            {
                std::string q = seq_of(g, level_strains[si].path);
                LD d = after[q] - before[q];
                if (abundance[si] < Zt || d < 0.01 * dmax) drop.push_back(si);
            }
            for (auto it = drop.rbegin(); it != drop.rend(); ++it) level_strains.erase(level_strains.begin() + *it);
        }
        else if (!level_reads.empty())
            hard_clustering(g, level_strains, level_reads, new_reads, pairs);

        // ---- candidate strains of the next level
        branching = false;
        for (auto& s : level_strains)
        {
            const int v = s.path.back();
            const Node& nv = g.store[v];
            LD oz = 0, moc = 0;
            std::vector<LD> oc;
            for (int o : nv.out)
            {
                int c = g.cover(v, o);
                oc.push_back(c);
                oz += c;
                if (moc < c) moc = c;
            }
            int dd = 0;
            for (size_t oi = 0; oi < nv.out.size(); ++oi)
            {
                const int o = nv.out[oi];
                if (g.store[o].label != "$" && oz > 0)
                {
                    if (oc[oi] <= 1. && oc[oi] < moc) { dd += 1; continue; }
                    Strain ns = s;
                    ns.path.push_back(o);
                    if (oc[oi] > 0) ns.abundance = s.abundance * oc[oi] / oz;
                    else ns.abundance = oz * std::min(0.01, (double)tau);
                    next_strains.push_back(ns);
                }
                else
                {
                    Strain ns = s;
                    ns.path.push_back(o);
                    ns.abundance = s.abundance;
                    next_strains.push_back(ns);
                }
            }
            if ((int)nv.out.size() > 1 + dd) branching = true;
        }
        if (next_strains.size() > 80)
        {
            std::vector<LD> ssa;
            for (auto& s : next_strains) ssa.push_back(s.abundance);
            std::sort(ssa.begin(), ssa.end(), [](LD x, LD y) { return x > y; });
            const LD cut = ssa[80];
            std::vector<Strain> keep;
            for (auto& s : next_strains) if (!(s.abundance < cut)) keep.push_back(s);
            next_strains.swap(keep);
        }

        while (!nxt.empty()) { cur.push(nxt.front()); nxt.pop(); }
        level_strains.swap(next_strains);
        next_strains.clear();
        queued.clear();
        level_reads.clear();
        level_read_count = 0;
    }
    (void)level_read_count;

    out.clear();
    for (auto& s : result)
    {
        StrainOut o;
        o.abundance = s.abundance;
        o.Z = s.Z;
        o.path = s.path;
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) o.sub[i][j] = s.sub[i][j];
        o.loglik = s.ll;
        out.push_back(o);
    }
}

// read_assign(strains, vector<AlignRead>, read_pairs, n), NonparametricClustering.cpp:776-836
void read_assign(const Pog& g, const std::vector<Read>& reads, const PairTable& pairs, int n,
                 std::vector<StrainOut>& strains)
{
    (void)g;
    int read_size = 0;
    for (auto& r : reads) read_size += r.cn;
    n = std::min(n, 40000 / read_size);
    const size_t S = strains.size();
    std::vector<LD> a(S), p(S, 0);
    for (size_t i = 0; i < S; ++i) a[i] = strains[i].abundance;
    Sampler sampler;
    for (; n > 0; n--)
        for (int id = 0; id < (int)reads.size(); ++id)
            for (int cn = reads[id].cn; cn > 0; --cn)
            {
                int uid = mate_of(pairs, id, cn - 1);
                p = a;
                normalize(p);
                for (size_t i = 0; i < S; ++i)
                {
                    p[i] = std::log(p[i]) + strains[i].loglik[id];
                    if (uid >= 0) p[i] += strains[i].loglik[uid];
                    p[i] = std::exp(p[i]);
                }
                int c = sampler.draw(p);
                a[c] += 1;
            }
    normalize(a);
    for (size_t i = 0; i < S; ++i) strains[i].abundance = a[i];
}

}  // namespace oracle
