// Host graph construction for the B200 StrainCall engine (see pog.hpp for the phase split).
//
// Behavioural contract: the node order, states, labels, ordered successor / predecessor lists and
// ordered read pools equal those of the reference's PartialOrderGraph(G,R)
// (/root/reference/StrainCall/PartialOrderGraph.cpp:61-265 and the helpers it calls, cited per
// function below).  What differs is how the work is done:
//   * nodes keep stable handles; nothing is renumbered while nodes are dropped (the reference
//     shifts every later id on each delete_node, PartialOrderGraph.cpp:421-441) -- ids are handed
//     out once, at flatten time, in creation order of the survivors, which is the same order;
//   * the insertion alignments are not run level by level on the host: all levels are listed up
//     front (they only depend on the threaded reads) and solved in one device launch;
//   * the "level without deletions" of every node (PartialOrderGraph.cpp:571-622) is taken from ONE
//     traversal per graph, since deletion canonisation never changes the deletion-free subgraph;
//   * reads covering an edge are counted by a sorted merge (or a binary search from the shorter pool) instead of the
//     quadratic pool scan (PartialOrderGraph.cpp:1218-1244), once per edge;
//   * a read pool is an ascending list of read ids wherever its entries are implied by the node (one letter, the read's
//     copies) -- all but ~500 of the 650 000 entries of a 5 000-read subgroup -- and a list of (read, string, copies)
//     items only where merges made it one.
#include "pog.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iterator>
#include <map>
#include <queue>
#include <set>
#include <sstream>
#include <stack>

namespace rambl {

namespace {

// The letters of a read-pool entry: one letter nearly always (a few after path collapse), so they live inline -- a
// std::string per entry (32 bytes, 650 000 entries per 5 000-read subgroup) was the larger half of the builder's memory
// traffic.  Longer strings (long collapsed paths) go to the heap.
class PoolStr
{
public:
    PoolStr() : n_(0) { u_.inl[0] = 0; }
    explicit PoolStr(char c) : n_(1) { u_.inl[0] = c; }
    PoolStr(const char* p) : n_(0) { assign(p, (uint32_t)strlen(p)); }
    PoolStr(const PoolStr& o) : n_(0) { assign(o.data(), o.n_); }
    PoolStr(PoolStr&& o) noexcept : n_(o.n_), u_(o.u_) { o.n_ = 0; }
    PoolStr& operator=(const PoolStr& o) { if (this != &o) { release(); assign(o.data(), o.n_); } return *this; }
    PoolStr& operator=(PoolStr&& o) noexcept { if (this != &o) { release(); n_ = o.n_; u_ = o.u_; o.n_ = 0; } return *this; }
    ~PoolStr() { release(); }
    size_t size() const { return n_; }
    const char* data() const { return n_ <= kInline ? u_.inl : u_.heap; }
    char operator[](size_t i) const { return data()[i]; }
    int compare(const PoolStr& o) const
    {
        const int c = memcmp(data(), o.data(), std::min(n_, o.n_));
        return c != 0 ? c : (n_ < o.n_ ? -1 : (n_ > o.n_ ? 1 : 0));
    }
    friend PoolStr operator+(const PoolStr& a, const PoolStr& b)
    {
        PoolStr r;
        r.n_ = a.n_ + b.n_;
        char* d = r.n_ <= kInline ? r.u_.inl : (r.u_.heap = (char*)malloc(r.n_));
        memcpy(d, a.data(), a.n_);
        memcpy(d + a.n_, b.data(), b.n_);
        return r;
    }

private:
    static constexpr uint32_t kInline = 8;
    void assign(const char* p, uint32_t n)
    {
        n_ = n;
        char* d = n <= kInline ? u_.inl : (u_.heap = (char*)malloc(n));
        memcpy(d, p, n);
    }
    void release() { if (n_ > kInline) free(u_.heap); n_ = 0; }
    uint32_t n_;
    union { char inl[8]; char* heap; } u_;
};

struct PoolItem
{
    int rid;
    PoolStr s;
    int cn;
    bool operator<(const PoolItem& o) const
    {
        if (rid != o.rid) return rid < o.rid;
        const int c = s.compare(o.s);
        if (c != 0) return c < 0;
        return cn < o.cn;
    }
};

// A vertex's read pool has one of two forms.  Nearly every entry of a graph belongs to a match / mismatch node, which
// a read enters at most once, with the node's own letter and the read's copy number: such a pool is just the ascending
// list of read ids (`rids`, 4 bytes per entry instead of a 32-byte PoolItem) -- the "plain" form.  Anything else
// (per-read insertion / deletion nodes, merged or collapsed nodes) keeps explicit items in `pool`.  A plain pool is
// turned into items the moment something needs them (Impl::itemise).
struct Vertex
{
    uint8_t st = ST_MAT;
    bool alive = true;
    bool plain = false;     // the pool is `rids` x (`letter`, copies of the read); `pool` is empty
    char letter = 0;
    std::string label;
    int level = -1;
    std::vector<int> out, in, sib;
    std::vector<int> rids;
    std::vector<PoolItem> pool;
};

struct Span  // a gap chain hanging between two aligned nodes
{
    int from, to;
    std::vector<int> chain;
};

struct LevelPlan  // insertion spans of one backbone level, already in alignment order
{
    int level;
    std::vector<Span> spans;
    std::vector<std::string> seqs;
    int problem = -1;  // index into the MSA batch, or -1 when all spans have one length
    int width = 0;     // common width when problem == -1
};

void remove_first(std::vector<int>& v, int x)
{
    for (size_t i = 0; i < v.size(); ++i)
        if (v[i] == x) { v.erase(v.begin() + i); return; }
}

}  // namespace

struct GraphBuilder::Impl
{
    std::vector<Vertex> V;
    int backbone = 0;  // handles 0..backbone+1 are ^, gene letters, $
    std::vector<LevelPlan> plans;
    int first_problem = 0, n_problems = 0;
    int n_reads = 0;
    std::vector<int> copies;  // per read id (every pool entry of a read carries the read's copy number)
    // Guard against inputs on which the reference's construction never ends (alignments that make the
    // graph cyclic, e.g. reads that begin with a deletion): every loop below pays into one budget.
    mutable long long work = 0;
    void tick(const char* where) const
    {
        if (++work > 100000000LL + 200LL * (long long)V.size())
            throw Error(RAMBL_ERR_INVALID, std::string("graph construction does not terminate (malformed alignments?) in ") + where);
    }

    int add(uint8_t st, const std::string& label)
    {
        V.emplace_back();
        V.back().st = st;
        V.back().label = label;
        return (int)V.size() - 1;
    }
    int add(uint8_t st, char c) { return add(st, std::string(1, c)); }
    // a node whose pool will be in the plain form
    int add_plain(uint8_t st, char c)
    {
        const int h = add(st, std::string(1, c));
        V[h].plain = true;
        V[h].letter = c;
        return h;
    }
    size_t pool_size(int h) const { return V[h].plain ? V[h].rids.size() : V[h].pool.size(); }
    int pool_front(int h) const { return V[h].plain ? V[h].rids[0] : V[h].pool[0].rid; }  // read id of the first entry
    void itemise(int h)
    {
        Vertex& x = V[h];
        if (!x.plain) return;
        x.pool.reserve(x.rids.size());
        for (int rid : x.rids) x.pool.push_back({rid, PoolStr(x.letter), copies[rid]});
        std::vector<int>().swap(x.rids);
        x.plain = false;
    }
    bool has_edge(int u, int v) const
    {
        for (int o : V[u].out) if (o == v) return true;
        return false;
    }
    void connect(int u, int v) { V[u].out.push_back(v); V[v].in.push_back(u); }
    void disconnect(int u, int v) { remove_first(V[u].out, v); remove_first(V[v].in, u); }
    void hang(int u, const std::vector<int>& chain)
    {
        for (int g : chain) { connect(u, g); u = g; }
    }
    void hang(int u, int v, const std::vector<int>& chain)
    {
        hang(u, chain);
        connect(chain.back(), v);
    }

    // delete_node, PartialOrderGraph.cpp:406-444 (no renumbering here, see file header)
    void drop(int w, bool bridge)
    {
        const std::vector<int> in = V[w].in, out = V[w].out;
        for (int x : in)
        {
            for (int y : out)
            {
                if (bridge && !has_edge(x, y)) connect(x, y);
                remove_first(V[y].in, w);
            }
            remove_first(V[x].out, w);
        }
        V[w].alive = false;
    }

    // merge_node + merge_read_pool, PartialOrderGraph.cpp:963-1038
    void absorb(int u, int v)
    {
        {
            const std::vector<int> vin = V[v].in;
            for (int x : vin) if (x != u && !has_edge(x, u)) connect(x, u);
            const std::vector<int> vout = V[v].out;
            for (int y : vout) if (y != u && !has_edge(u, y)) connect(u, y);
        }
        const bool chained = has_edge(u, v) && V[u].st == ST_MAT && V[v].st == ST_MAT;
        if (chained) V[u].label += V[v].label;
        if (!chained && V[u].plain && V[v].plain && V[u].letter == V[v].letter)
        {
            // two plain pools with the same letter and no read in common stay plain: the union of the ids
            std::vector<int>& ra = V[u].rids;
            const std::vector<int>& rb = V[v].rids;
            std::vector<int> r(ra.size() + rb.size());
            const auto end = std::set_union(ra.begin(), ra.end(), rb.begin(), rb.end(), r.begin());
            if ((size_t)(end - r.begin()) == ra.size() + rb.size())
            {
                ra.swap(r);
                drop(v, false);
                return;
            }
        }
        itemise(u);
        itemise(v);
        std::vector<PoolItem>& a = V[u].pool;
        std::vector<PoolItem>& b = V[v].pool;
        std::sort(a.begin(), a.end());
        std::sort(b.begin(), b.end());
        std::vector<PoolItem> r;
        r.reserve(a.size() + b.size());
        size_t i = 0, j = 0;
        while (i < a.size() && j < b.size())
        {
            if (a[i].rid == b[j].rid) { r.push_back({a[i].rid, a[i].s + b[j].s, a[i].cn}); ++i; ++j; }
            else if (a[i].rid < b[j].rid) r.push_back(a[i++]);
            else r.push_back(b[j++]);
        }
        for (; i < a.size(); ++i) r.push_back(a[i]);
        for (; j < b.size(); ++j) r.push_back(b[j]);
        a.swap(r);
        drop(v, false);
    }

    // ---- phase A -------------------------------------------------------------------------------
    // PartialOrderGraph::build up to (not including) canonize_graph, PartialOrderGraph.cpp:67-255
    void splice_reads(const std::string& G, const ReadSet& R)
    {
        V.clear();
        V.reserve(G.size() + 2 + R.size());
        backbone = (int)G.size();
        n_reads = (int)R.size();
        copies.resize(R.size());
        for (size_t rid = 0; rid < R.size(); ++rid) copies[rid] = R.cn[rid];
        int u = add(ST_MAT, '^');
        for (char c : G) { int w = add_plain(ST_MAT, c); connect(u, w); u = w; }
        connect(u, add(ST_MAT, '$'));
        const int last = backbone + 1;
        {
            // backbone pools grow to the coverage of their position: size them once (difference array over the
            // reads' reference spans) instead of doubling their way up
            std::vector<int> cover(G.size() + 2, 0);
            for (size_t rid = 0; rid < R.size(); ++rid)
            {
                const int pos = R.pos[rid];
                if (pos < 0 || pos >= last) continue;  // refused below
                const int hi = (int)std::min<size_t>(G.size(), (size_t)pos + R.seq(rid).size());
                ++cover[pos];
                --cover[hi];
            }
            int depth = 0;
            for (size_t i = 0; i < G.size(); ++i)
            {
                depth += cover[i];
                if (depth > 0) V[i + 1].rids.reserve((size_t)depth);
            }
        }

        for (int rid = 0; rid < (int)R.size(); ++rid)
        {
            const int pos = R.pos[rid];
            if (pos < 0 || pos >= last) throw Error(RAMBL_ERR_INVALID, "read starts outside the gene window");
            const std::string_view r = R.seq(rid);
            int i = pos, j = 0;
            u = pos;
            int v = pos + 1;
            auto step_v = [&]() {
                ++i;
                if (i + 1 > last) throw Error(RAMBL_ERR_INVALID, "read runs past the end of the gene window");
                v = i + 1;
            };
            size_t k = 0;
            const std::string_view cg = R.cigar(rid);
            while (k < cg.size())
            {
                int len = 0;
                bool digits = false;
                while (k < cg.size() && cg[k] >= '0' && cg[k] <= '9') { len = len * 10 + (cg[k] - '0'); ++k; digits = true; }
                if (k >= cg.size() || !digits) throw Error(RAMBL_ERR_INVALID, "malformed CIGAR");
                char op = cg[k++];
                if (op == '=' || op == 'X') op = 'M';  // parse_cigar, PartialOrderGraph.cpp:46-53
                if (op == 'S') { j += j + len; continue; }  // sic, PartialOrderGraph.cpp:126
                if (op == 'M')
                {
                    if (j + len > (int)r.size()) throw Error(RAMBL_ERR_INVALID, "CIGAR longer than the read");
                    for (int q = 0; q < len; ++q, ++j)
                    {
                        if (v > last || i >= backbone) throw Error(RAMBL_ERR_INVALID, "read runs past the end of the gene window");
                        const uint8_t st = (G[i] == r[j]) ? ST_MAT : ST_MIS;
                        int hit = -1;
                        if (V[v].st == st && V[v].label.size() == 1 && V[v].label[0] == r[j]) hit = v;
                        else
                            for (int s : V[v].sib)
                                if (V[s].st == st && V[s].label[0] == r[j]) { hit = s; break; }
                        if (hit < 0)
                        {
                            hit = add_plain(st, r[j]);
                            connect(u, hit);
                            V[v].sib.push_back(hit);
                        }
                        else if (!(hit == v && u == v - 1) && !has_edge(u, hit)) connect(u, hit);  // backbone edges are there from the start
                        V[hit].rids.push_back(rid);  // a read passes a column once: ids stay ascending
                        u = hit;
                        step_v();
                    }
                }
                else if (op == 'I')
                {
                    if (j + len > (int)r.size()) throw Error(RAMBL_ERR_INVALID, "CIGAR longer than the read");
                    std::vector<int> chain;
                    for (int q = 0; q < len; ++q, ++j)
                    {
                        int w = add_plain(ST_INS, r[j]);
                        V[w].rids.push_back(rid);
                        chain.push_back(w);
                    }
                    if (chain.empty()) continue;
                    hang(u, chain);
                    u = chain.back();
                }
                else if (op == 'D')
                {
                    std::vector<int> chain;
                    for (int q = 0; q < len; ++q)
                    {
                        int w = add_plain(ST_DEL, '=');
                        V[w].rids.push_back(rid);
                        chain.push_back(w);
                        step_v();
                    }
                    if (chain.empty()) continue;
                    if (v == last) { hang(u, v, chain); u = v; continue; }
                    hang(u, chain);
                    u = chain.back();
                }
                // N, H, P: carried by parse_cigar but ignored by build
            }
            if (u != v && !has_edge(u, v)) connect(u, v);
        }
    }

    // find_insert_from, PartialOrderGraph.cpp:355-393 (the chain buffer is shared by all branches, as there)
    void spans_from(int u, std::vector<Span>& out) const
    {
        std::vector<int> chain;
        std::vector<int> todo(1, u);
        while (!todo.empty())
        {
            tick("insertion scan");
            const int v = todo.back();
            todo.pop_back();
            const Vertex& nv = V[v];
            if (v == u)
            {
                for (int o : nv.out) if (V[o].st == ST_INS) todo.push_back(o);
            }
            else if (nv.st == ST_MAT || nv.st == ST_MIS)
            {
                out.push_back({u, v, chain});
                chain.clear();
            }
            else
            {
                chain.push_back(v);
                for (int o : nv.out) todo.push_back(o);
            }
        }
    }

    void plan_insertions(MsaBatch& batch)
    {
        plans.clear();
        first_problem = batch.problems();
        n_problems = 0;
        for (int i = 0; i <= backbone; ++i)
        {
            LevelPlan lp;
            lp.level = i;
            spans_from(i, lp.spans);
            for (int s : V[i].sib) spans_from(s, lp.spans);
            if (lp.spans.empty()) continue;
            // same comparator, same sequence as PartialOrderGraph.cpp:466 -> same order, ties included
            std::sort(lp.spans.begin(), lp.spans.end(), [](Span& a, Span& b) { return a.chain.size() > b.chain.size(); });
            size_t lmax = 0, lmin = (size_t)-1;
            for (const Span& sp : lp.spans)
            {
                std::string s;
                for (int h : sp.chain) s += V[h].label;
                lmax = std::max(lmax, s.size());
                lmin = std::min(lmin, s.size());
                lp.seqs.push_back(s);
            }
            lp.width = (int)lmax;
            if (lp.spans.size() > 1 && lmax != lmin)
            {
                lp.problem = first_problem + n_problems++;
                for (const std::string& s : lp.seqs) batch.add_sequence(s.data(), (int)s.size());
                batch.end_problem();
            }
            plans.push_back(std::move(lp));
        }
    }

    // ---- phase C -------------------------------------------------------------------------------
    // find_common_read_pool, PartialOrderGraph.cpp:780-829 -> (rid, cn) in set order
    std::vector<std::pair<int, int>> shared_reads(int a, int b) const
    {
        typedef std::vector<std::pair<int, int>> Pairs;
        auto list = [&](int h, Pairs& to) {
            const Vertex& x = V[h];
            if (x.plain) for (int rid : x.rids) to.push_back({rid, copies[rid]});
            else for (const PoolItem& p : x.pool) to.push_back({p.rid, p.cn});
        };
        // the (rid, cn) set of node h without those of its insertion / deletion neighbours on one side, ascending
        auto own = [&](int h, const std::vector<int>& nb) {
            Pairs all, gone, left;
            list(h, all);
            for (int o : nb)
                if (V[o].st == ST_INS || V[o].st == ST_DEL) list(o, gone);
            std::sort(all.begin(), all.end());
            all.erase(std::unique(all.begin(), all.end()), all.end());
            if (gone.empty()) return all;
            std::sort(gone.begin(), gone.end());
            std::set_difference(all.begin(), all.end(), gone.begin(), gone.end(), std::back_inserter(left));
            return left;
        };
        const Pairs ra = own(a, V[a].out), rb = own(b, V[b].in);
        Pairs c;
        std::set_intersection(ra.begin(), ra.end(), rb.begin(), rb.end(), std::back_inserter(c));
        return c;
    }

    // the (u | u.sibling) x (v | v.sibling) pairings of add_edge(int,int) / delete_edge(int)
    std::vector<std::pair<int, int>> pairings(int i) const
    {
        std::vector<std::pair<int, int>> pr;
        const int u = i, v = i + 1;
        pr.push_back({u, v});
        for (int s : V[v].sib) pr.push_back({u, s});
        for (int s : V[u].sib) pr.push_back({s, v});
        for (int su : V[u].sib) for (int sv : V[v].sib) pr.push_back({su, sv});
        return pr;
    }

    // canonize_insert_at_level with the aligned rows supplied, PartialOrderGraph.cpp:446-550,831-961
    void settle_insertions(const LevelPlan& lp, const MsaResult& rows)
    {
        int width = lp.width;
        if (lp.problem >= 0)
        {
            for (size_t t = 0; t < lp.spans.size(); ++t)
            {
                const std::string aligned = rows.row(lp.problem, (int)t);
                if (aligned == lp.seqs[t]) continue;
                int rid = 0;
                for (int h : lp.spans[t].chain)
                {
                    rid = pool_front(h);
                    drop(h, true);
                }
                std::vector<int> chain;
                for (char c : aligned)
                {
                    int w = add_plain(ST_INS, c);
                    V[w].rids.push_back(rid);
                    chain.push_back(w);
                }
                hang(lp.spans[t].from, lp.spans[t].to, chain);
            }
            width = rows.width[lp.problem];
        }
        const std::vector<std::pair<int, int>> pr = pairings(lp.level);
        for (const auto& e : pr)
        {
            if (!has_edge(e.first, e.second)) continue;
            const std::vector<std::pair<int, int>> crp = shared_reads(e.first, e.second);
            std::vector<int> chain;
            for (int t = 0; t < width; ++t)
            {
                // the shared reads come in ascending id order, one entry per read, with the read's copies: a plain pool
                int w = add_plain(ST_INS, '-');
                V[w].rids.reserve(crp.size());
                for (const auto& r : crp) V[w].rids.push_back(r.first);
                chain.push_back(w);
            }
            hang(e.first, e.second, chain);
        }
        for (const auto& e : pr)
            if (has_edge(e.first, e.second)) disconnect(e.first, e.second);
    }

    // one traversal instead of node_level_exclude_delete per query (PartialOrderGraph.cpp:571-622)
    void deletion_free_levels(std::vector<int>& first_seen, int& exhausted) const
    {
        first_seen.assign(V.size(), -1);
        std::vector<int> cur(1, 0), nxt;
        std::vector<int> mark(V.size(), -1);
        int level = 0;
        while (!cur.empty())
        {
            tick("deletion-free levels");
            const int u = cur.back();
            cur.pop_back();
            if (first_seen[u] < 0) first_seen[u] = level;
            for (int o : V[u].out)
            {
                if (V[o].st == ST_DEL) continue;
                nxt.push_back(o);
                for (int s : V[o].sib) nxt.push_back(s);
            }
            if (cur.empty())
            {
                while (!nxt.empty())
                {
                    const int v = nxt.back();
                    nxt.pop_back();
                    if (mark[v] == level) continue;
                    mark[v] = level;
                    cur.push_back(v);
                }
                level += 1;
            }
        }
        exhausted = level;
    }

    // find_delete_from, PartialOrderGraph.cpp:624-672
    void deletions_from(int w, std::vector<Span>& out) const
    {
        std::vector<int> chain;
        std::vector<std::pair<int, int>> todo(1, {w, 0});
        while (!todo.empty())
        {
            tick("deletion scan");
            const int u = todo.back().first, pass = todo.back().second;
            todo.pop_back();
            const Vertex& nu = V[u];
            if (u == w)
            {
                for (int o : nu.out) if (V[o].st == ST_DEL) todo.push_back({o, 0});
            }
            else if (nu.st == ST_DEL)
            {
                if (pass == 0)
                {
                    todo.push_back({u, 1});
                    chain.push_back(u);
                    for (int o : nu.out) todo.push_back({o, 0});
                }
                else chain.pop_back();
            }
            else out.push_back({w, u, chain});
        }
    }

    // canonize_delete, PartialOrderGraph.cpp:684-752
    void settle_deletions()
    {
        std::vector<int> lvl;
        int exhausted = 0;
        deletion_free_levels(lvl, exhausted);
        auto level_of = [&](int h) { return lvl[h] >= 0 ? lvl[h] : exhausted; };
        for (int i = 0; i <= backbone; ++i)
        {
            std::vector<Span> dels;
            deletions_from(i, dels);
            const std::vector<int> sibs = V[i].sib;
            for (int s : sibs) deletions_from(s, dels);
            for (const Span& d : dels)
            {
                const int want = level_of(d.to) - level_of(d.from) - 1;
                const int have = (int)d.chain.size();
                if (want - have <= 0) continue;
                const int first = d.chain[0];
                const int rid = pool_front(first);
                std::vector<int> chain;
                for (int t = want - have; t > 0; --t)
                {
                    int w = add_plain(ST_DEL, '=');
                    V[w].rids.push_back(rid);
                    chain.push_back(w);
                }
                hang(d.from, first, chain);
                disconnect(d.from, first);
            }
        }
    }

    // forward_merge / backward_merge, PartialOrderGraph.cpp:1040-1159
    void merge_equal_neighbours(bool forward)
    {
        std::queue<int> todo;
        std::vector<char> seen(V.size(), 0), merged(V.size(), 0);
        if (forward) todo.push(0);
        else
            for (int h = 0; h < (int)V.size(); ++h)
                if (V[h].alive && V[h].label == "$") todo.push(h);
        std::vector<std::pair<int, int>> plan;
        while (!todo.empty())
        {
            tick("sibling merge");
            const int w = todo.front();
            todo.pop();
            if (merged[w]) continue;
            plan.clear();
            {
                const std::vector<int>& adj = forward ? V[w].out : V[w].in;
                for (size_t a = 0; a < adj.size(); ++a)
                    for (size_t b = a + 1; b < adj.size(); ++b)
                    {
                        const int u = adj[a], v = adj[b];
                        if (u == v) continue;
                        if (V[u].st != V[v].st || V[u].label != V[v].label) continue;
                        if (merged[u] || merged[v]) continue;
                        plan.push_back({u, v});
                        merged[v] = 1;
                    }
            }
            for (const auto& pr : plan) absorb(pr.first, pr.second);
            const std::vector<int> adj = forward ? V[w].out : V[w].in;
            for (int x : adj)
                if (!seen[x]) { todo.push(x); seen[x] = 1; }
        }
    }

    // path_collapse, PartialOrderGraph.cpp:1171-1216
    void collapse_chains()
    {
        size_t level_size = 0;
        std::vector<int> cur(1, 0), nxt;
        std::vector<int> mark(V.size(), -1);
        int level = 0;
        size_t head = 0;
        while (head < cur.size())
        {
            tick("path collapse");
            const int u = cur[head++];
            if (level_size == 1 && V[u].out.size() == 1)
            {
                int v = V[u].out[0];
                while (V[v].out.size() == 1)
                {
                    tick("path collapse");
                    absorb(u, v);
                    v = V[u].out[0];
                }
            }
            for (int v : V[u].out)
                if (mark[v] != level) { mark[v] = level; nxt.push_back(v); }
            if (head == cur.size())
            {
                cur.swap(nxt);
                nxt.clear();
                head = 0;
                level += 1;
                level_size = cur.size();
            }
        }
    }

    // node_level() via LevelOrderIterator, LevelOrderIterator.cpp:3-56
    void level_nodes()
    {
        int N = 0;
        for (const Vertex& x : V) N += x.alive ? 1 : 0;
        std::vector<int> cur, nxt;
        std::vector<int> mark(V.size(), -1);
        int epoch = 0;
        int n = 0, level = 0, at = 0, at_level = 0;
        for (int o : V[0].out) cur.push_back(o);
        mark[0] = epoch;
        while (n != N)
        {
            tick("node levels");
            V[at].level = at_level;
            if (nxt.empty()) { level += 1; ++epoch; }
            if (!cur.empty())
            {
                const int w = cur.back();
                cur.pop_back();
                at = w;
                at_level = level;
                n += 1;
                for (int o : V[w].out) nxt.push_back(o);
                if (cur.empty())
                    while (!nxt.empty())
                    {
                        const int x = nxt.back();
                        nxt.pop_back();
                        if (mark[x] == epoch) continue;
                        mark[x] = epoch;
                        cur.push_back(x);
                    }
            }
            else n += 1;
        }
    }

    void flatten(FlatGraph& g) const
    {
        std::vector<int> id(V.size(), -1);
        int N = 0;
        size_t n_pool = 0, n_pool_chars = 0, n_out = 0, n_in = 0, n_label = 0;
        std::vector<char> in_read_order(V.size(), 1);  // pool ids ascending (almost always: reads are threaded in order)
        for (size_t h = 0; h < V.size(); ++h)
        {
            const Vertex& x = V[h];
            if (!x.alive) continue;
            id[h] = N++;
            n_out += x.out.size(); n_in += x.in.size(); n_label += x.label.size();
            if (x.plain) { n_pool += x.rids.size(); n_pool_chars += x.rids.size(); continue; }
            n_pool += x.pool.size();
            for (size_t k = 0; k < x.pool.size(); ++k)
            {
                n_pool_chars += x.pool[k].s.size();
                if (k && x.pool[k].rid < x.pool[k - 1].rid) in_read_order[h] = 0;
            }
        }
        g = FlatGraph();
        g.n_nodes = N;
        g.n_reads = n_reads;
        g.st.reserve(N); g.level.reserve(N);
        g.label_chars.reserve(n_label); g.label_off.reserve(N + 1);
        g.out_to.reserve(n_out); g.out_cover.reserve(n_out); g.out_off.reserve(N + 1);
        g.in_from.reserve(n_in); g.in_off.reserve(N + 1);
        g.pool_rid.resize(n_pool); g.pool_cn.resize(n_pool); g.pool_off.reserve(N + 1);
        g.pool_chars.resize(n_pool_chars); g.pool_str_off.resize(n_pool + 1);
        g.pool_str_off[0] = 0;
        g.label_off.assign(1, 0); g.out_off.assign(1, 0); g.in_off.assign(1, 0); g.pool_off.assign(1, 0);
        size_t at_pool = 0, at_chars = 0;
        std::vector<int> ids;  // the read ids of the current vertex, ascending (when its pool is not plain)
        for (size_t h = 0; h < V.size(); ++h)
        {
            const Vertex& x = V[h];
            if (!x.alive) continue;
            g.st.push_back(x.st);
            g.level.push_back(x.level);
            g.label_chars.insert(g.label_chars.end(), x.label.begin(), x.label.end());
            g.label_off.push_back((int)g.label_chars.size());
            if (x.label == "$") g.end_node = id[h];
            const int* a = x.rids.data();
            size_t na = x.rids.size();
            if (!x.out.empty() && !x.plain)
            {
                ids.clear();
                for (const PoolItem& p : x.pool) ids.push_back(p.rid);
                if (!in_read_order[h]) std::sort(ids.begin(), ids.end());
                a = ids.data();
                na = ids.size();
            }
            for (int o : x.out)
            {
                g.out_to.push_back(id[o]);
                g.out_cover.push_back(reads_over_edge((int)h, o, a, na, in_read_order[o] != 0));
            }
            g.out_off.push_back((int)g.out_to.size());
            for (int o : x.in) g.in_from.push_back(id[o]);
            g.in_off.push_back((int)g.in_from.size());
            if (x.plain)
            {
                const size_t n = x.rids.size();
                if (n)
                {
                    memcpy(g.pool_rid.data() + at_pool, x.rids.data(), n * sizeof(int));
                    memset(g.pool_chars.data() + at_chars, x.letter, n);
                }
                int* cn = g.pool_cn.data() + at_pool;
                int* so = g.pool_str_off.data() + at_pool + 1;
                for (size_t k = 0; k < n; ++k) { cn[k] = copies[x.rids[k]]; so[k] = (int)(at_chars + k + 1); }
                at_pool += n;
                at_chars += n;
            }
            else
                for (const PoolItem& p : x.pool)
                {
                    g.pool_rid[at_pool] = p.rid;
                    g.pool_cn[at_pool] = p.cn;
                    const size_t len = p.s.size();
                    if (len == 1) g.pool_chars[at_chars] = p.s[0];
                    else memcpy(g.pool_chars.data() + at_chars, p.s.data(), len);
                    at_chars += len;
                    g.pool_str_off[++at_pool] = (int)at_chars;
                }
            g.pool_off.push_back((int)at_pool);
        }
    }

    // number_of_reads_cover_nodes, PartialOrderGraph.cpp:1218-1244: sum over pairs with equal read id of
    // the second pool's copy number.  `a[0..na)` holds the first pool's ids in ascending order.  Two plain pools are
    // intersected from the shorter side when one is much shorter (a mismatch node next to a backbone node), else merged
    // in one pass; a second pool with explicit items is merged if it is in read order and looked up item by item if not.
    int reads_over_edge(int hu, int hv, const int* a, size_t na, bool v_in_read_order) const
    {
        const Vertex& u = V[hu];
        const Vertex& v = V[hv];
        int n = 0;
        if (hu == 0 || v.label == "$")
        {
            const Vertex& w = hu == 0 ? v : u;
            if (w.plain) for (int rid : w.rids) n += copies[rid];
            else for (const PoolItem& p : w.pool) n += p.cn;
            return n;
        }
        if (v.plain)
        {
            const int* b = v.rids.data();
            const size_t nb = v.rids.size();
            if (nb * 8 < na)
            {
                for (size_t k = 0; k < nb; ++k)
                {
                    auto range = std::equal_range(a, a + na, b[k]);
                    n += (int)(range.second - range.first) * copies[b[k]];
                }
                return n;
            }
            if (na * 8 < nb)
            {
                for (size_t k = 0; k < na; ++k)  // repeated ids in a: each occurrence counts
                    if (std::binary_search(b, b + nb, a[k])) n += copies[a[k]];
                return n;
            }
            size_t i = 0;
            for (size_t k = 0; k < nb; ++k)
            {
                const int rid = b[k];
                while (i < na && a[i] < rid) ++i;
                size_t j = i;
                while (j < na && a[j] == rid) ++j;
                n += (int)(j - i) * copies[rid];
                i = j;  // ids of a plain pool are distinct
            }
            return n;
        }
        if (v_in_read_order)
        {
            size_t i = 0;
            for (const PoolItem& p : v.pool)
            {
                while (i < na && a[i] < p.rid) ++i;
                size_t j = i;
                while (j < na && a[j] == p.rid) ++j;  // i stays: the next item of v may carry the same id
                n += (int)(j - i) * p.cn;
            }
            return n;
        }
        for (const PoolItem& p : v.pool)
        {
            auto range = std::equal_range(a, a + na, p.rid);
            n += (int)(range.second - range.first) * p.cn;
        }
        return n;
    }
};

GraphBuilder::GraphBuilder() : m(new Impl) {}
GraphBuilder::~GraphBuilder() { delete m; }
int GraphBuilder::n_problems() const { return m->n_problems; }
void GraphBuilder::rebase_problems(int first)
{
    const int shift = first - m->first_problem;
    for (LevelPlan& lp : m->plans) if (lp.problem >= 0) lp.problem += shift;
    m->first_problem = first;
}

void GraphBuilder::thread(const std::string& gene, const ReadSet& reads, MsaBatch& batch)
{
    m->splice_reads(gene, reads);
    m->plan_insertions(batch);
}

void GraphBuilder::finish(const MsaResult& rows, FlatGraph& out)
{
    if (m->n_problems > 0 && (int)rows.width.size() < m->first_problem + m->n_problems)
        throw Error(RAMBL_ERR_STATE, "GraphBuilder::finish called without the solved alignment batch");
    static const bool trace = getenv("RAMBL_TRACE") != nullptr;
    auto t = std::chrono::steady_clock::now();
    double ms[7];
    int k = 0;
    auto lap = [&] {
        const auto n = std::chrono::steady_clock::now();
        ms[k++] = std::chrono::duration<double, std::milli>(n - t).count();
        t = n;
    };
    for (const LevelPlan& lp : m->plans) m->settle_insertions(lp, rows);
    lap();
    m->settle_deletions();
    lap();
    m->merge_equal_neighbours(true);
    m->merge_equal_neighbours(false);
    lap();
    m->collapse_chains();
    lap();
    m->level_nodes();
    lap();
    m->flatten(out);
    lap();
    if (trace)
        fprintf(stderr, "[rambl] graph finish ms: insertions %.1f deletions %.1f merge %.1f collapse %.1f level %.1f flatten %.1f\n",
                ms[0], ms[1], ms[2], ms[3], ms[4], ms[5]);
}

void fill_edge_cover(FlatGraph& g)
{
    g.out_cover.assign(g.out_to.size(), 0);
    std::vector<int> a;
    for (int u = 0; u < g.n_nodes; ++u)
    {
        a.assign(g.pool_rid.begin() + g.pool_off[u], g.pool_rid.begin() + g.pool_off[u + 1]);
        std::sort(a.begin(), a.end());
        int own = 0;
        for (int e = g.pool_off[u]; e < g.pool_off[u + 1]; ++e) own += g.pool_cn[e];
        for (int e = g.out_off[u]; e < g.out_off[u + 1]; ++e)
        {
            const int v = g.out_to[e];
            int n = 0;
            if (u == 0) { for (int q = g.pool_off[v]; q < g.pool_off[v + 1]; ++q) n += g.pool_cn[q]; }
            else if (v == g.end_node) n = own;
            else
                for (int q = g.pool_off[v]; q < g.pool_off[v + 1]; ++q)
                {
                    auto range = std::equal_range(a.begin(), a.end(), g.pool_rid[q]);
                    n += (int)(range.second - range.first) * g.pool_cn[q];
                }
            g.out_cover[e] = n;
        }
    }
}

std::string FlatGraph::dump() const
{
    std::ostringstream os;
    os << "NODES " << n_nodes << "\n";
    for (int u = 0; u < n_nodes; ++u)
    {
        os << "NODE " << u << " " << (int)st[u] << " " << label(u) << " " << level[u] << " | OUT";
        for (int e = out_off[u]; e < out_off[u + 1]; ++e) os << " " << out_to[e];
        os << " | IN";
        for (int e = in_off[u]; e < in_off[u + 1]; ++e) os << " " << in_from[e];
        os << " | SIB | POOL";
        for (int e = pool_off[u]; e < pool_off[u + 1]; ++e) os << " " << pool_rid[e] << ":" << pool_str(e) << ":" << pool_cn[e];
        os << "\n";
    }
    return os.str();
}

std::string FlatGraph::edges() const
{
    std::ostringstream os;
    for (int u = 0; u < n_nodes; ++u)
    {
        int rc = 0;
        for (int e = pool_off[u]; e < pool_off[u + 1]; ++e) rc += pool_cn[e];
        os << "#\t" << u << "\t" << level[u] << "\t" << label(u) << "\t" << rc << "\n";
    }
    for (int u = 0; u < n_nodes; ++u)
        for (int e = out_off[u]; e < out_off[u + 1]; ++e) os << u << "\t" << out_to[e] << "\t" << out_cover[e] << "\n";
    return os.str();
}

}  // namespace rambl
