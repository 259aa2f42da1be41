// oracle/oracle_pog.cpp -- TEST INFRASTRUCTURE ONLY (CPU restatement; never linked into the product).
//
// Restatement of the reference's partial order graph construction,
//   PartialOrderGraph::build            /root/reference/StrainCall/PartialOrderGraph.cpp:67-265
//   canonize_insert(_at_level)          PartialOrderGraph.cpp:355-403,446-564,831-961
//   canonize_delete(_at_level)          PartialOrderGraph.cpp:571-752
//   forward_merge / backward_merge      PartialOrderGraph.cpp:963-1159
//   path_collapse                       PartialOrderGraph.cpp:1171-1216
//   node_level (LevelOrderIterator)     PartialOrderGraph.cpp:769-776, LevelOrderIterator.cpp:3-56
//   output_edge / reads covering edge   PartialOrderGraph.cpp:318-337,1218-1244
// Nodes live in a handle-indexed store; `order` is the reference's `nodes` vector.
// Parity status: PINNED against oracle/_ref by tests/test_oracle_vs_ref.py (node order, states,
// levels, ordered out/in/sibling lists and ordered read pools must all be identical).
//
// Where the reference would run into undefined behaviour (erasing set::end() in
// find_common_read_pool, PartialOrderGraph.cpp:795-796,812-813) the restatement skips the erase.
#include "oracle.h"

#include <algorithm>
#include <queue>
#include <set>
#include <sstream>
#include <stack>
#include <tuple>

namespace oracle {

namespace {
struct Cig { char op; int len; };

// parse_cigar, PartialOrderGraph.cpp:13-59 ('=' and 'X' read as 'M')
std::vector<Cig> split_cigar(const std::string& c)
{
    std::vector<Cig> r;
    std::string num;
    for (char ch : c)
    {
        switch (ch)
        {
            case 'M': case 'I': case 'D': case 'N': case 'S': case 'H': case 'P':
                r.push_back({ch, std::stoi(num)}); num.clear(); break;
            case '=': case 'X':
                r.push_back({'M', std::stoi(num)}); num.clear(); break;
            default: num.push_back(ch);
        }
    }
    return r;
}

void erase_first(std::vector<int>& v, int x)
{
    for (auto it = v.begin(); it != v.end(); ++it)
        if (*it == x) { v.erase(it); return; }
}

struct GapEx { int u, v; std::vector<int> gap; };
typedef std::set<std::pair<int, int>> RidSet;  // (rid, cn)
}  // namespace

int Pog::new_node(int st, const std::string& label)
{
    Node n;
    n.id = (int)order.size();
    n.st = st;
    n.label = label;
    store.push_back(n);
    order.push_back((int)store.size() - 1);
    return (int)store.size() - 1;
}
void Pog::link(int u, int w) { store[u].out.push_back(w); store[w].in.push_back(u); }
void Pog::unlink(int u, int v) { erase_first(store[u].out, v); erase_first(store[v].in, u); }
bool Pog::linked(int u, int v) const
{
    for (int o : store[u].out) if (o == v) return true;
    return false;
}
void Pog::link_chain(int u, const std::vector<int>& gap)
{
    int a = u;
    for (int g : gap) { link(a, g); a = g; }
}
void Pog::link_chain(int u, int v, const std::vector<int>& gap)
{
    link_chain(u, gap);
    link(gap.back(), v);
}

// delete_node, PartialOrderGraph.cpp:406-444
void Pog::drop_node(int w, bool bridging)
{
    // note: iterates w's own lists, which only other nodes' edits touch
    const std::vector<int> in = store[w].in, out = store[w].out;
    for (int x : in)
    {
        for (int y : out)
        {
            if (bridging && !linked(x, y)) link(x, y);
            erase_first(store[y].in, w);
        }
        erase_first(store[x].out, w);
    }
    const int wid = store[w].id;
    if (wid >= 0 && wid < (int)order.size() && order[wid] == w) order.erase(order.begin() + wid);
    for (int i = wid; i < (int)order.size(); ++i) store[order[i]].id -= 1;
}

// merge_read_pool + merge_node, PartialOrderGraph.cpp:963-1038
void Pog::fuse(int u, int v)
{
    {
        const std::vector<int> vin = store[v].in;
        for (int x : vin) if (!linked(x, u) && x != u) link(x, u);
        const std::vector<int> vout = store[v].out;
        for (int y : vout) if (!linked(u, y) && u != y) link(u, y);
    }
    if (linked(u, v) && store[u].st == ST_MAT && store[v].st == ST_MAT) store[u].label += store[v].label;

    std::vector<PoolEntry>& a = store[u].pool;
    std::vector<PoolEntry>& b = store[v].pool;
    std::sort(a.begin(), a.end());
    std::sort(b.begin(), b.end());
    std::vector<PoolEntry> r;
    size_t i = 0, j = 0;
    while (i < a.size() && j < b.size())
    {
        if (a[i].rid == b[j].rid) { r.push_back({a[i].rid, a[i].s + b[j].s, a[i].cn}); ++i; ++j; }
        else if (a[i].rid < b[j].rid) r.push_back(a[i++]);
        else r.push_back(b[j++]);
    }
    while (i < a.size()) r.push_back(a[i++]);
    while (j < b.size()) r.push_back(b[j++]);
    a = r;
    drop_node(v, false);
}

// number_of_reads_cover_nodes, PartialOrderGraph.cpp:1218-1244
int Pog::cover(int hu, int hv) const
{
    const Node& u = store[hu];
    const Node& v = store[hv];
    int n = 0;
    if (hu == order[0]) { for (auto& p : v.pool) n += p.cn; }
    else if (v.label == "$") { for (auto& p : u.pool) n += p.cn; }
    else
        for (auto& a : u.pool)
            for (auto& b : v.pool)
                if (a.rid == b.rid) n += b.cn;
    return n;
}

namespace {

// find_insert_from, PartialOrderGraph.cpp:355-393
void inserts_from(const Pog& g, int u, std::vector<GapEx>& out)
{
    std::vector<int> gap;
    std::stack<int> todo;
    todo.push(u);
    while (!todo.empty())
    {
        int v = todo.top(); todo.pop();
        const Node& nv = g.store[v];
        if (v == u)
        {
            for (int o : nv.out) if (g.store[o].st == ST_INS) todo.push(o);
        }
        else if (nv.st == ST_MAT || nv.st == ST_MIS)
        {
            out.push_back({u, v, gap});
            gap.clear();
        }
        else
        {
            gap.push_back(v);
            for (int o : nv.out) todo.push(o);
        }
    }
}

// find_delete_from, PartialOrderGraph.cpp:624-672
void deletes_from(const Pog& g, int w, std::vector<GapEx>& out)
{
    std::vector<int> gap;
    std::stack<std::pair<int, int>> todo;
    todo.push({w, 0});
    while (!todo.empty())
    {
        int u = todo.top().first, c = todo.top().second;
        todo.pop();
        const Node& nu = g.store[u];
        if (u == w)
        {
            for (int o : nu.out) if (g.store[o].st == ST_DEL) todo.push({o, 0});
        }
        else if (nu.st == ST_DEL)
        {
            if (c == 0)
            {
                todo.push({u, 1});
                gap.push_back(u);
                for (int o : nu.out) todo.push({o, 0});
            }
            else gap.pop_back();
        }
        else out.push_back({w, u, gap});
    }
}

// find_common_read_pool, PartialOrderGraph.cpp:780-829
RidSet common_reads(const Pog& g, int a, int b)
{
    RidSet ra, rb, c;
    for (auto& p : g.store[a].pool) ra.insert({p.rid, p.cn});
    for (int o : g.store[a].out)
        if (g.store[o].st == ST_INS || g.store[o].st == ST_DEL)
            for (auto& p : g.store[o].pool) { auto it = ra.find({p.rid, p.cn}); if (it != ra.end()) ra.erase(it); }
    for (auto& p : g.store[b].pool) rb.insert({p.rid, p.cn});
    for (int o : g.store[b].in)
        if (g.store[o].st == ST_INS || g.store[o].st == ST_DEL)
            for (auto& p : g.store[o].pool) { auto it = rb.find({p.rid, p.cn}); if (it != rb.end()) rb.erase(it); }
    for (auto& x : ra) if (rb.count(x)) c.insert(x);
    return c;
}

// the four (u|u.sibling) x (v|v.sibling) pairings of add_edge(int,int)/delete_edge(int), in the reference's order
std::vector<std::pair<int, int>> level_pairs(const Pog& g, int i)
{
    int u = g.order[i], v = g.order[i + 1];
    std::vector<std::pair<int, int>> pr;
    pr.push_back({u, v});
    for (int s : g.store[v].sib) pr.push_back({u, s});
    for (int s : g.store[u].sib) pr.push_back({s, v});
    for (int su : g.store[u].sib) for (int sv : g.store[v].sib) pr.push_back({su, sv});
    return pr;
}

// add_edge(int i,int l), PartialOrderGraph.cpp:831-923
void pad_level(Pog& g, int i, int l)
{
    // the sibling lists do not change here, so the pairing can be listed up front
    for (auto& pr : level_pairs(g, i))
    {
        if (!g.linked(pr.first, pr.second)) continue;
        RidSet crp = common_reads(g, pr.first, pr.second);
        std::vector<int> gap;
        for (int t = 0; t < l; ++t)
        {
            int w = g.new_node(ST_INS, "-");
            for (auto& r : crp) g.store[w].pool.push_back({r.first, "-", r.second});
            gap.push_back(w);
        }
        g.link_chain(pr.first, pr.second, gap);
    }
}
// delete_edge(int i), PartialOrderGraph.cpp:925-961
void cut_level(Pog& g, int i)
{
    for (auto& pr : level_pairs(g, i))
        if (g.linked(pr.first, pr.second)) g.unlink(pr.first, pr.second);
}

// canonize_insert_at_level, PartialOrderGraph.cpp:446-550
void canon_insert_level(Pog& g, int i)
{
    std::vector<GapEx> ins;
    {
        int u = g.order[i];
        inserts_from(g, u, ins);
        const std::vector<int> sibs = g.store[u].sib;
        for (int s : sibs) inserts_from(g, s, ins);
    }
    if (ins.empty()) return;
    std::sort(ins.begin(), ins.end(), [](GapEx& a, GapEx& b) { return a.gap.size() > b.gap.size(); });
    std::vector<std::string> seqs;
    size_t lmax = 0, lmin = 1000000000;
    for (auto& x : ins)
    {
        std::string s;
        for (int h : x.gap) s += g.store[h].label;
        seqs.push_back(s);
        lmax = std::max(lmax, s.size());
        lmin = std::min(lmin, s.size());
    }
    const int n = (int)ins.size();
    int l = (int)lmax;
    if (n > 1 && lmax != lmin)
    {
        std::vector<std::string> rows = msa_sp_align(seqs);
        for (int t = 0; t < n; ++t)
        {
            if (rows[t] == seqs[t]) continue;
            int rid = 0, rcn = 0;
            for (int h : ins[t].gap)
            {
                rid = g.store[h].pool[0].rid;
                rcn = g.store[h].pool[0].cn;
                g.drop_node(h, true);
            }
            std::vector<int> gap;
            for (char c : rows[t])
            {
                int w = g.new_node(ST_INS, std::string(1, c));
                g.store[w].pool.push_back({rid, std::string(1, c), rcn});
                gap.push_back(w);
            }
            g.link_chain(ins[t].u, ins[t].v, gap);
        }
        l = rows.empty() ? 0 : (int)rows[0].size();
    }
    pad_level(g, i, l);
    cut_level(g, i);
}

// node_level_exclude_delete, PartialOrderGraph.cpp:571-622
int level_skipping_deletes(const Pog& g, int w)
{
    std::stack<int> cur, nxt;
    std::set<int> seen;
    int level = 0;
    cur.push(g.order[0]);
    while (!cur.empty())
    {
        int u = cur.top(); cur.pop();
        if (u == w) break;
        for (int o : g.store[u].out)
        {
            if (g.store[o].st == ST_DEL) continue;
            nxt.push(o);
            for (int s : g.store[o].sib) nxt.push(s);
        }
        if (cur.empty())
        {
            while (!nxt.empty())
            {
                int v = nxt.top(); nxt.pop();
                if (seen.count(v)) continue;
                cur.push(v);
                seen.insert(v);
            }
            level += 1;
            seen.clear();
        }
    }
    return level;
}

// canonize_delete_at_level, PartialOrderGraph.cpp:684-740
void canon_delete_level(Pog& g, int i)
{
    std::vector<GapEx> dels;
    {
        int u = g.order[i];
        deletes_from(g, u, dels);
        const std::vector<int> sibs = g.store[u].sib;
        for (int s : sibs) deletes_from(g, s, dels);
    }
    if (dels.empty()) return;
    std::map<int, int> lvl;
    for (auto& d : dels)
    {
        int u = d.u, v = d.v;
        if (!lvl.count(u)) lvl[u] = level_skipping_deletes(g, u);
        if (!lvl.count(v)) lvl[v] = level_skipping_deletes(g, v);
        int want = lvl[v] - lvl[u] - 1;
        int have = (int)d.gap.size();
        if (want - have > 0)
        {
            int first = d.gap[0];
            int rid = g.store[first].pool[0].rid, rcn = g.store[first].pool[0].cn;
            std::vector<int> gap;
            for (int t = want - have; t > 0; --t)
            {
                int w = g.new_node(ST_DEL, "=");
                g.store[w].pool.push_back({rid, "=", rcn});
                gap.push_back(w);
            }
            g.link_chain(u, first, gap);
            g.unlink(u, first);
        }
    }
}

// forward_merge / backward_merge, PartialOrderGraph.cpp:1040-1159
void sweep_merge(Pog& g, bool forward)
{
    std::queue<int> todo;
    std::set<int> seen, merged;
    if (forward) todo.push(g.order[0]);
    else
        for (int h : g.order) if (g.store[h].label == "$") todo.push(h);
    while (!todo.empty())
    {
        int w = todo.front(); todo.pop();
        if (merged.count(w)) continue;
        std::vector<std::pair<int, int>> plan;
        {
            const std::vector<int>& adj = forward ? g.store[w].out : g.store[w].in;
            for (size_t a = 0; a < adj.size(); ++a)
                for (size_t b = a + 1; b < adj.size(); ++b)
                {
                    int u = adj[a], v = adj[b];
                    if (u == v) continue;
                    if (g.store[u].st == g.store[v].st && g.store[u].label == g.store[v].label)
                        if (!merged.count(u) && !merged.count(v)) { plan.push_back({u, v}); merged.insert(v); }
                }
        }
        for (auto& p : plan) g.fuse(p.first, p.second);
        const std::vector<int> adj = forward ? g.store[w].out : g.store[w].in;
        for (int x : adj) if (!seen.count(x)) { todo.push(x); seen.insert(x); }
    }
}

// path_collapse, PartialOrderGraph.cpp:1171-1216
void collapse_paths(Pog& g)
{
    size_t level_size = 0;
    std::queue<int> cur, nxt;
    std::set<int> queued;
    cur.push(g.order[0]);
    while (!cur.empty())
    {
        int u = cur.front(); cur.pop();
        if (level_size == 1 && g.store[u].out.size() == 1)
        {
            int v = g.store[u].out[0];
            while (g.store[v].out.size() == 1)
            {
                g.fuse(u, v);
                v = g.store[u].out[0];
            }
        }
        for (int v : g.store[u].out) if (!queued.count(v)) { nxt.push(v); queued.insert(v); }
        if (cur.empty())
        {
            while (!nxt.empty()) { cur.push(nxt.front()); nxt.pop(); }
            level_size = cur.size();
            queued.clear();
        }
    }
}

// node_level() through LevelOrderIterator, LevelOrderIterator.cpp:3-56
void assign_levels(Pog& g)
{
    const int N = (int)g.order.size();
    std::stack<int> cur, nxt;
    std::set<int> seen;
    int n = 0, level = 0;
    int at = g.order[0], at_level = 0;
    for (int o : g.store[at].out) cur.push(o);
    seen.insert(at);
    while (n != N)
    {
        g.store[at].level = at_level;
        // operator++
        if (nxt.empty()) { level += 1; seen.clear(); }
        if (!cur.empty())
        {
            int w = cur.top(); cur.pop();
            at = w;
            at_level = level;
            n += 1;
            for (int o : g.store[w].out) nxt.push(o);
            if (cur.empty())
                while (!nxt.empty())
                {
                    int x = nxt.top(); nxt.pop();
                    if (seen.count(x)) continue;
                    cur.push(x);
                    seen.insert(x);
                }
        }
        else n += 1;
    }
}
}  // namespace

void Pog::build(const std::string& G, const std::vector<Read>& R)
{
    store.clear();
    order.clear();
    int B = new_node(ST_MAT, "^");
    int u = B;
    for (char c : G) { int w = new_node(ST_MAT, std::string(1, c)); link(u, w); u = w; }
    int E = new_node(ST_MAT, "$");
    link(u, E);

    for (int rid = 0; rid < (int)R.size(); ++rid)
    {
        const Read& rd = R[rid];
        int i = rd.pos, j = 0;
        u = order[rd.pos];
        int v = order[rd.pos + 1];
        const std::string& r = rd.seq;
        for (const Cig& c : split_cigar(rd.cigar))
        {
            if (c.op == 'S') { j += j + c.len; continue; }  // sic, PartialOrderGraph.cpp:126
            if (c.op == 'M')
            {
                for (int k = 0; k < c.len; ++k, ++j)
                {
                    const int st = (G[i] == r[j]) ? ST_MAT : ST_MIS;
                    const std::string lab(1, r[j]);
                    int hit = -1;
                    if (store[v].st == st && store[v].label == lab) hit = v;
                    else
                        for (int s : store[v].sib)
                            if (store[s].st == st && store[s].label == lab) { hit = s; break; }
                    if (hit < 0)
                    {
                        int w = new_node(st, lab);
                        link(u, w);
                        store[w].pool.push_back({rid, lab, rd.cn});
                        store[v].sib.push_back(w);
                        u = w;
                    }
                    else
                    {
                        if (!linked(u, hit)) link(u, hit);
                        store[hit].pool.push_back({rid, lab, rd.cn});
                        u = hit;
                    }
                    v = order[++i + 1];
                }
            }
            else if (c.op == 'I')
            {
                std::vector<int> gap;
                for (int k = 0; k < c.len; ++k, ++j)
                {
                    int w = new_node(ST_INS, std::string(1, r[j]));
                    store[w].pool.push_back({rid, std::string(1, r[j]), rd.cn});
                    gap.push_back(w);
                }
                link_chain(u, gap);
                u = gap.back();
            }
            else if (c.op == 'D')
            {
                std::vector<int> gap;
                for (int k = 0; k < c.len; ++k)
                {
                    int w = new_node(ST_DEL, "=");
                    store[w].pool.push_back({rid, "=", rd.cn});
                    gap.push_back(w);
                    v = order[++i + 1];
                }
                if (store[v].label == "$") { link_chain(u, v, gap); u = v; continue; }
                link_chain(u, gap);
                u = gap.back();
            }
        }
        if (!linked(u, v) && u != v) link(u, v);
    }

    // canonize_graph, PartialOrderGraph.cpp:754-767
    for (int i = 0; i < (int)order.size(); ++i)
    {
        if (store[order[i]].label == "$") break;
        canon_insert_level(*this, i);
    }
    for (int i = 0; i < (int)order.size(); ++i)
    {
        if (store[order[i]].label == "$") break;
        canon_delete_level(*this, i);
    }
    sweep_merge(*this, true);
    sweep_merge(*this, false);
    collapse_paths(*this);
    assign_levels(*this);
}

std::string Pog::dump() const
{
    std::ostringstream os;
    os << "NODES " << order.size() << "\n";
    for (int h : order)
    {
        const Node& u = store[h];
        os << "NODE " << u.id << " " << u.st << " " << u.label << " " << u.level;
        os << " | OUT"; for (int o : u.out) os << " " << store[o].id;
        os << " | IN";  for (int o : u.in) os << " " << store[o].id;
        os << " | SIB"; for (int o : u.sib) os << " " << store[o].id;
        os << " | POOL"; for (auto& p : u.pool) os << " " << p.rid << ":" << p.s << ":" << p.cn;
        os << "\n";
    }
    return os.str();
}

std::string Pog::edges() const
{
    std::ostringstream os;
    for (int h : order)
    {
        const Node& u = store[h];
        int rc = 0;
        for (auto& p : u.pool) rc += p.cn;
        os << "#\t" << u.id << "\t" << u.level << "\t" << u.label << "\t" << rc << "\n";
    }
    for (int h : order)
        for (int o : store[h].out) os << store[h].id << "\t" << store[o].id << "\t" << cover(h, o) << "\n";
    return os.str();
}

}  // namespace oracle
