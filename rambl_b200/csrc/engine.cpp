// Host orchestration of the level-synchronous strain search (see engine.hpp).
#include "engine.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <random>
#include <unordered_map>
#include <chrono>
#include <cstdlib>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "dpm.cuh"
#include "walk.cuh"

namespace rambl {

namespace {

// 80 candidates plus their children before the cut; a subgroup that needs more gets private buffers
constexpr int kInitialSlots = 192;
constexpr int kUniforms = 40000;  // a call never draws more: sweeps = min(n, 40000 / reads)

// std::uniform draws of std::discrete_distribution on std::mt19937(1234): generate_canonical<double,53>
// takes two 32-bit outputs, low word first (libstdc++ bits/random.tcc:3349-3381).
std::vector<double> canonical_stream(int n)
{
    std::mt19937 gen(1234);
    std::vector<double> u(n);
    for (int i = 0; i < n; ++i)
    {
        const double lo = (double)gen(), hi = (double)gen();
        double x = (lo + hi * 4294967296.0) / 18446744073709551616.0;
        if (x >= 1.0) x = std::nextafter(1.0, 0.0);
        u[i] = x;
    }
    return u;
}

// per-group scratch in the weights buffer (layout in dpm.cu): [D padded to 32 / 32][S][32] weight tiles + [D]
// normalisers + [D] letter codes, kept 256-byte aligned
inline long long scratch_doubles(int S, int D)
{
    const long long Dp = (D + 31) & ~31;
    return ((long long)S * Dp + 2LL * D + 31) & ~31LL;
}

struct Cand  // a candidate strain on the host: everything per-read lives in its device slot
{
    int slot = -1;
    double ab = 0;
    int tail = -1;   // index into Sub::trail
    int node = -1;   // last node of the path
    uint64_t hash = 1469598103934665603ull;  // of the concatenated labels (the reference keys maps by strain_seq)
    uint64_t len = 0;
};

struct Sub
{
    const SubgroupInput* in = nullptr;
    const FlatGraph* g = nullptr;
    int R = 0;
    // device state: carved out of the engine's arenas at start (one allocation for the whole batch);
    // only a subgroup that outgrows its slots gets private buffers
    struct DevPtr { double* p = nullptr; } ll, sub;  // [slot][read] log-likelihoods, [slot][36] model counts
    const char* d_label = nullptr;
    const char* d_pool = nullptr;
    DevBuf<double> ll_own, sub_own;
    int slot_cap = 0;
    std::vector<int> free_slots;
    std::vector<char> retained;
    // walk state
    std::vector<Cand> cands;
    std::vector<std::pair<int, int>> trail;  // (parent trail index, node)
    std::vector<int> cur, nxt, mark;
    int epoch = 0;
    bool branching = false, done = false, failed = false;
    std::vector<uint8_t> present;
    int levels = 0;
    // this step: every subgroup stages its share privately (the host part of a level runs on all cores),
    // the shares are then packed into one pinned arena and go to the device in one copy
    int mode = MODE_NONE, m = 0, D = 0, read_size = 0, nsweeps = 0;
    int ab_off = 0;
    std::vector<int> I;
    std::vector<double> Dv;
    StepGroup sg;
    bool has_group = false;
    std::vector<std::pair<int, int>> ops;  // (src slot, dst slot) copies to run before the next level
    int local_launches = 0;
    long long step_updates = 0, step_draws = 0;
    int side = -1;  // solved by the side batch (level-synchronous path, own stream and thread): index there
    // result
    bool have_result = false;
    std::vector<Cand> result;
    int status = RAMBL_OK;
    long long draws = 0;
};

inline uint64_t extend_hash(uint64_t h, const char* s, int n)
{
    for (int i = 0; i < n; ++i) { h ^= (unsigned char)s[i]; h *= 1099511628211ull; }
    return h;
}

std::vector<int> path_of(const Sub& s, int tail)
{
    std::vector<int> p;
    for (int t = tail; t >= 0; t = s.trail[t].first) p.push_back(s.trail[t].second);
    std::reverse(p.begin(), p.end());
    return p;
}

// seq_identity, NonparametricClustering.cpp:584-612
double sequence_identity(const std::string& a, const std::string& b)
{
    int iden = 0, len = 0;
    for (size_t i = 0; i < a.size(); ++i)
    {
        const char x = a[i], y = i < b.size() ? b[i] : '\0';
        if ((x == '-' || x == '=') && (y == '-' || y == '=')) continue;
        if (x == '^' && y == '^') continue;
        if (x == y) iden += 1;
        len += 1;
    }
    return (iden + 0.0) / len;
}

// std::sort with the reference's comparator on the same sequence: the same permutation, ties included
template <typename T>
void sort_desc_by_abundance(std::vector<T>& v)
{
    std::vector<int> idx(v.size());
    for (size_t i = 0; i < idx.size(); ++i) idx[i] = (int)i;
    std::sort(idx.begin(), idx.end(), [&](int a, int b) { return v[a].ab > v[b].ab; });
    std::vector<T> r;
    r.reserve(v.size());
    for (int i : idx) r.push_back(v[i]);
    v.swap(r);
}

// Persistent workers for the per-subgroup host work of a level (a few microseconds each, thousands of levels).
class Workers
{
public:
    explicit Workers(unsigned n) : stop_(false), gen_(0), pending_(0)
    {
        for (unsigned t = 0; t < n; ++t) th_.emplace_back([this] { loop(); });
    }
    ~Workers()
    {
        {
            std::unique_lock<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    // fn(i) for i in [0, n); returns when all are done; the first exception is rethrown
    void run(size_t n, const std::function<void(size_t)>& fn)
    {
        if (n == 0) return;
        if (th_.empty() || n == 1)
        {
            for (size_t i = 0; i < n; ++i) fn(i);
            return;
        }
        {
            std::unique_lock<std::mutex> lk(mu_);
            fn_ = &fn;
            n_ = n;
            next_.store(0);
            pending_ = th_.size();
            err_ = nullptr;
            ++gen_;
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
        if (err_) std::rethrow_exception(err_);
    }

private:
    void loop()
    {
        unsigned long seen = 0;
        for (;;)
        {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
            }
            for (;;)
            {
                const size_t i = next_.fetch_add(1);
                if (i >= n_) break;
                try { (*fn_)(i); }
                catch (...)
                {
                    std::unique_lock<std::mutex> lk(mu_);
                    if (!err_) err_ = std::current_exception();
                }
            }
            std::unique_lock<std::mutex> lk(mu_);
            if (--pending_ == 0) done_.notify_all();
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    bool stop_;
    unsigned long gen_;
    size_t pending_, n_ = 0;
    std::atomic<size_t> next_{0};
    const std::function<void(size_t)>* fn_ = nullptr;
    std::exception_ptr err_;
};

struct Engine
{
    const InferParams& prm;
    cudaStream_t st;
    EngineStats& stats;
    std::vector<Sub> subs;
    double tau, diff, e;
    // step staging
    std::vector<StepGroup> h_groups;
    PinBuf<char> p_arena;   // the step's staged inputs: [doubles | step groups | ints]
    size_t off_groups = 0, off_I = 0, arena_bytes = 0;
    PinBuf<double> p_al;
    size_t n_I = 0, n_D = 0;
    std::vector<int> group_sub;
    std::unique_ptr<Workers> workers;
    DevBuf<char> d_arena;
    DevBuf<double> d_W, d_U;
    DevBuf<InheritOp> d_ops;
    DevBuf<unsigned long long> d_counters;
    DevBuf<double> ll_arena, sub_arena;
    DevBuf<char> char_arena, pool_arena;
    long long w_total = 0, max_stride = 0;
    std::vector<cudaEvent_t> gibbs_events;  // pairs, resolved after the last step

    Engine(const InferParams& p, cudaStream_t s, EngineStats& es) : prm(p), st(s), stats(es)
    {
        tau = (double)p.tau;
        diff = (double)p.diff;
        e = (double)p.e;
    }

    // ---- slots ---------------------------------------------------------------------------------
    void grow_slots(Sub& s, int want)
    {
        if (want <= s.slot_cap) return;
        int cap = std::max(s.slot_cap, 8);
        while (cap < want) cap *= 2;
        DevBuf<double> nll, nsub;
        nll.reserve((size_t)cap * s.R);
        nsub.reserve((size_t)cap * 36);
        RAMBL_CUDA(cudaMemsetAsync(nll.p, 0, sizeof(double) * (size_t)cap * s.R, st));
        launch_init_models(nsub.p, cap, e, st, &s.local_launches);
        RAMBL_CUDA(cudaMemcpyAsync(nll.p, s.ll.p, sizeof(double) * (size_t)s.slot_cap * s.R, cudaMemcpyDeviceToDevice, st));
        RAMBL_CUDA(cudaMemcpyAsync(nsub.p, s.sub.p, sizeof(double) * (size_t)s.slot_cap * 36, cudaMemcpyDeviceToDevice, st));
        RAMBL_CUDA(cudaStreamSynchronize(st));  // the old private buffers (if any) are freed below
        s.ll_own.swap(nll);
        s.sub_own.swap(nsub);
        s.ll.p = s.ll_own.p;
        s.sub.p = s.sub_own.p;
        for (int k = cap - 1; k >= s.slot_cap; --k) s.free_slots.push_back(k);
        s.retained.resize(cap, 0);
        s.slot_cap = cap;
    }
    int take_slot(Sub& s)
    {
        if (s.free_slots.empty()) grow_slots(s, s.slot_cap + 1);
        const int k = s.free_slots.back();
        s.free_slots.pop_back();
        return k;
    }
    void give_slot(Sub& s, int k)
    {
        if (k >= 0 && !s.retained[k]) s.free_slots.push_back(k);
    }
    void inherit(Sub& s, int src, int dst)
    {
        // grow_slots may move the buffers: the copies are resolved to pointers when they are launched
        s.ops.push_back({src, dst});
    }

    // ---- set-up --------------------------------------------------------------------------------
    void start(const std::vector<SubgroupInput>& in)
    {
        subs.resize(in.size());
        std::vector<double> u = canonical_stream(kUniforms);
        d_U.reserve(kUniforms);
        d_counters.reserve(4);
        RAMBL_CUDA(cudaMemsetAsync(d_counters.p, 0, 4 * sizeof(unsigned long long), st));
        RAMBL_CUDA(cudaMemcpyAsync(d_U.p, u.data(), sizeof(double) * kUniforms, cudaMemcpyHostToDevice, st));
        for (size_t i = 0; i < in.size(); ++i)
        {
            Sub& s = subs[i];
            s.in = &in[i];
            s.g = in[i].graph;
            if (!s.g) throw Error(RAMBL_ERR_INVALID, "subgroup without a graph");
            s.R = std::max(1, s.g->n_reads);
            if ((int)in[i].read_cn.size() != s.g->n_reads || (int)in[i].pair_off.size() != s.g->n_reads + 1)
                throw Error(RAMBL_ERR_INVALID, "read tables do not match the graph");
            for (int r = 0; r < s.g->n_reads; ++r)
                if (in[i].pair_off[r + 1] - in[i].pair_off[r] < in[i].read_cn[r])
                    throw Error(RAMBL_ERR_INVALID, "ReadPairs needs one entry per read copy");
            max_stride = std::max<long long>(max_stride, s.R);
            s.present.assign(s.R, 0);
            s.slot_cap = kInitialSlots;
            for (int k = kInitialSlots - 1; k >= 0; --k) s.free_slots.push_back(k);
            s.retained.assign(kInitialSlots, 0);
            s.mark.assign(s.g->n_nodes, -1);
            s.cur.assign(1, 0);
            if (s.g->n_nodes == 0) s.done = true;
        }
        // one allocation per kind for the whole batch (thousands of cudaMalloc/cudaFree pairs cost seconds)
        size_t n_ll = 0, n_sub = 0, n_ch = 0;
        for (const Sub& s : subs)
        {
            n_ll += (size_t)kInitialSlots * s.R;
            n_sub += (size_t)kInitialSlots * 36;
            n_ch += (s.g->label_chars.size() + 15) & ~(size_t)15;
        }
        ll_arena.reserve(std::max<size_t>(n_ll, 1));
        sub_arena.reserve(std::max<size_t>(n_sub, 1));
        char_arena.reserve(std::max<size_t>(n_ch, 16));
        RAMBL_CUDA(cudaMemsetAsync(ll_arena.p, 0, sizeof(double) * n_ll, st));
        launch_init_models(sub_arena.p, (int)(n_sub / 36), e, st, &stats.launches);
        // node labels now; the read-pool strings only for the subgroups the level-synchronous path ends up solving
        // (upload_pool_chars) -- the device walk has them in its own level tables
        std::vector<char> host_chars(std::max<size_t>(n_ch, 16), 0);
        size_t o_ll = 0, o_sub = 0, o_ch = 0;
        for (Sub& s : subs)
        {
            s.ll.p = ll_arena.p + o_ll;
            s.sub.p = sub_arena.p + o_sub;
            o_ll += (size_t)kInitialSlots * s.R;
            o_sub += (size_t)kInitialSlots * 36;
            s.d_label = char_arena.p + o_ch;
            if (!s.g->label_chars.empty()) memcpy(&host_chars[o_ch], s.g->label_chars.data(), s.g->label_chars.size());
            o_ch += (s.g->label_chars.size() + 15) & ~(size_t)15;
            s.d_pool = nullptr;
        }
        RAMBL_CUDA(cudaMemcpyAsync(char_arena.p, host_chars.data(), host_chars.size(), cudaMemcpyHostToDevice, st));
        RAMBL_CUDA(cudaStreamSynchronize(st));  // host_chars goes out of scope
        stats.h2d_bytes += (long long)host_chars.size();
        for (Sub& s : subs)
        {
            Cand root;  // Strain(100,e), NonparametricClustering.cpp:281
            root.slot = take_slot(s);
            s.cands.push_back(root);
        }
    }

    // the read-pool strings of the subgroups that still have levels to walk on the level-synchronous path
    void upload_pool_chars()
    {
        size_t n = 0;
        for (const Sub& s : subs) if (!s.done && !s.d_pool) n += (s.g->pool_chars.size() + 15) & ~(size_t)15;
        if (n == 0) return;
        pool_arena.reserve(n);
        std::vector<char> host(n, 0);
        size_t o = 0;
        for (Sub& s : subs)
        {
            if (s.done || s.d_pool) continue;
            s.d_pool = pool_arena.p + o;
            if (!s.g->pool_chars.empty()) memcpy(&host[o], s.g->pool_chars.data(), s.g->pool_chars.size());
            o += (s.g->pool_chars.size() + 15) & ~(size_t)15;
        }
        RAMBL_CUDA(cudaMemcpyAsync(pool_arena.p, host.data(), n, cudaMemcpyHostToDevice, st));
        RAMBL_CUDA(cudaStreamSynchronize(st));
        stats.h2d_bytes += (long long)n;
    }

    // ---- "$": read_reassign's sort + merge_strains, NonparametricClustering.cpp:309-315,645-702 ----
    void close_result(Sub& s, bool copy_out = true)
    {
        if (s.cands.empty()) { s.status = RAMBL_ERR_NO_STRAINS; s.have_result = true; s.result.clear(); return; }
        std::vector<Cand> v;
        v.swap(s.cands);
        sort_desc_by_abundance(v);
        sort_desc_by_abundance(v);  // merge_strains sorts again
        std::vector<Cand> merged(1, v[0]);
        std::vector<std::string> mseq(1, strain_sequence(*s.g, path_of(s, v[0].tail)));
        for (size_t i = 1; i < v.size(); ++i)
        {
            const std::string q = strain_sequence(*s.g, path_of(s, v[i].tail));
            size_t j = 0;
            for (; j < merged.size(); ++j)
                if (sequence_identity(q, mseq[j]) > 1 - diff) { merged[j].ab += v[i].ab; break; }
            if (j == merged.size()) { merged.push_back(v[i]); mseq.push_back(q); }
            else give_slot(s, v[i].slot);  // merged away: the level goes on without it
        }
        // the reference sorts and merges level_strains in place (lines 309-315): any level after this one
        // continues from the merged set
        s.cands = merged;
        // ... and copies the strains out; later levels (if any) must not touch the copies
        for (Cand& c : s.result) if (c.slot >= 0) { s.retained[c.slot] = 0; s.free_slots.push_back(c.slot); }
        if (copy_out)
            for (Cand& c : merged)
            {
                const int k = take_slot(s);
                inherit(s, c.slot, k);
                c.slot = k;
                s.retained[k] = 1;
            }
        s.result = merged;
        s.have_result = true;
        s.status = RAMBL_OK;
    }

    // ---- one level: host part before the launches ---------------------------------------------------
    void prepare(Sub& s)
    {
        s.I.clear();
        s.Dv.clear();
        s.has_group = false;
        s.step_updates = s.step_draws = 0;
        std::vector<int>& h_I = s.I;
        std::vector<double>& h_D = s.Dv;
        const FlatGraph& g = *s.g;
        std::vector<int> lv_rid, lv_so, lv_sl, lv_cn;
        s.nxt.clear();
        for (int u : s.cur)
        {
            if (u == 0) { s.cands[0].tail = (int)s.trail.size(); s.trail.push_back({-1, 0}); s.cands[0].node = 0;
                          s.cands[0].hash = extend_hash(s.cands[0].hash, g.label_chars.data() + g.label_off[0], g.label_off[1] - g.label_off[0]);
                          s.cands[0].ab = 1; }
            else if (g.label_off[u + 1] - g.label_off[u] == 1 && g.label_chars[g.label_off[u]] == '$') close_result(s);
            else
                for (int e2 = g.pool_off[u]; e2 < g.pool_off[u + 1]; ++e2)
                {
                    lv_rid.push_back(g.pool_rid[e2]);
                    lv_so.push_back(g.pool_str_off[e2]);
                    lv_sl.push_back(g.pool_str_off[e2 + 1] - g.pool_str_off[e2]);
                    lv_cn.push_back(g.pool_cn[e2]);
                }
            for (int e2 = g.out_off[u]; e2 < g.out_off[u + 1]; ++e2)
            {
                const int v = g.out_to[e2];
                if (s.mark[v] != s.epoch) { s.mark[v] = s.epoch; s.nxt.push_back(v); }
            }
        }
        const int m = (int)lv_rid.size(), S = (int)s.cands.size();
        s.m = m;
        s.mode = (m > 0 && S > 0) ? (s.branching ? MODE_GIBBS : MODE_HARD) : MODE_NONE;
        s.D = 0;
        s.read_size = 0;
        if (S > DPM_SMAX)
        {   // only seen when the abundances have degenerated (NaN survives the reference's "< cut" pruning)
            s.status = RAMBL_ERR_CAPACITY;
            s.have_result = false;
            s.failed = true;
            s.done = true;
            s.mode = MODE_NONE;
            return;
        }
        if (s.mode == MODE_NONE) return;

        StepGroup sg;
        memset(&sg, 0, sizeof sg);
        sg.ll = nullptr;  // resolved at launch
        sg.S = S;
        sg.m = m;
        sg.mode = s.mode;
        sg.slot_off = (int)h_I.size();
        for (const Cand& c : s.cands) h_I.push_back(c.slot);
        sg.lab_off = (int)h_I.size();
        bool any_multi = false;
        for (const Cand& c : s.cands) h_I.push_back(g.label_off[c.node]);
        for (const Cand& c : s.cands)
        {
            const int l = g.label_off[c.node + 1] - g.label_off[c.node];
            any_multi = any_multi || l > 1;
            h_I.push_back(l);
        }
        sg.rid_off = (int)h_I.size();
        h_I.insert(h_I.end(), lv_rid.begin(), lv_rid.end());
        h_I.insert(h_I.end(), lv_so.begin(), lv_so.end());
        h_I.insert(h_I.end(), lv_sl.begin(), lv_sl.end());
        // "new" = first time any strain sees the read (read_loglik.count(rid)==0, line 364); the flag
        // hard_clustering reads (new_reads, line 375) is only raised by strains on a collapsed node
        for (int r = 0; r < m; ++r)
        {
            const bool fresh = !s.present[lv_rid[r]];
            s.present[lv_rid[r]] = 1;
            h_I.push_back(fresh ? 1 : 0);
        }
        if (s.mode == MODE_HARD && !any_multi)
            for (int r = 0; r < m; ++r) h_I[sg.rid_off + 3 * m + r] = 0;
        // draws: one per read copy, in read order, copies counted down (lines 36-39, 169-189)
        int D = 0;
        for (int r = 0; r < m; ++r) D += lv_cn[r];
        sg.D = D;
        sg.read_size = D;
        sg.draw_off = (int)h_I.size();
        h_I.resize(h_I.size() + 2 * (size_t)D);
        int* dr = &h_I[sg.draw_off];
        int* dm = dr + D;
        int d = 0;
        const SubgroupInput& in = *s.in;
        for (int r = 0; r < m; ++r)
            for (int cn = lv_cn[r]; cn > 0; --cn, ++d)
            {
                const int rid = lv_rid[r];
                int mate = in.pair_val[in.pair_off[rid] + cn - 1];
                if (mate >= s.R) throw Error(RAMBL_ERR_INVALID, "mate id out of range");
                if (s.mode == MODE_GIBBS) { if (mate >= 0 && !s.present[mate]) mate = -1; }
                dr[d] = r;
                dm[d] = mate;
            }
        if (s.mode == MODE_HARD)  // Strain::logprob(uid) creates the entry (line 57)
            for (int k = 0; k < D; ++k) if (dm[k] >= 0) s.present[dm[k]] = 1;
        if (D <= 0) throw Error(RAMBL_ERR_INVALID, "a read-pool entry without copies");
        sg.nsweeps = (s.mode == MODE_GIBBS) ? std::min(prm.n, 40000 / D) : 0;
        sg.ab_off = (int)h_D.size();
        for (const Cand& c : s.cands) h_D.push_back(c.ab);
        s.D = D;
        s.read_size = D;
        s.nsweeps = sg.nsweeps;
        s.ab_off = sg.ab_off;
        s.sg = sg;
        s.has_group = true;
        s.step_updates = (long long)m * S;
        if (s.mode == MODE_GIBBS && S >= 2) { s.step_draws = (long long)D * sg.nsweeps; s.draws += s.step_draws; }
    }

    // ---- one level: host part after the launches ------------------------------------------------------
    void advance(Sub& s, const double* al)
    {
        const FlatGraph& g = *s.g;
        std::vector<Cand>& cs = s.cands;
        if (s.mode == MODE_GIBBS)
        {
            // abundance maps keyed by strain_seq (lines 404-429); equal sequences share an entry
            std::unordered_map<uint64_t, double> before, after;
            for (size_t i = 0; i < cs.size(); ++i) before[cs[i].hash ^ (cs[i].len * 0x9e3779b97f4a7c15ull)] = cs[i].ab;
            for (size_t i = 0; i < cs.size(); ++i) cs[i].ab += al[i];
            for (size_t i = 0; i < cs.size(); ++i) after[cs[i].hash ^ (cs[i].len * 0x9e3779b97f4a7c15ull)] = cs[i].ab;
            double dmax = 0, Z = 0;
            std::vector<double> delta(cs.size());
            for (size_t i = 0; i < cs.size(); ++i)
            {
                const uint64_t k = cs[i].hash ^ (cs[i].len * 0x9e3779b97f4a7c15ull);
                delta[i] = after[k] - before[k];
                if (dmax < delta[i]) dmax = delta[i];
                Z += al[i];
            }
            const double Zt = Z * tau;
            std::vector<Cand> keep;
            for (size_t i = 0; i < cs.size(); ++i)
            {
                if (al[i] < Zt || delta[i] < 0.01 * dmax) give_slot(s, cs[i].slot);
                else keep.push_back(cs[i]);
            }
            cs.swap(keep);
        }
        else if (s.mode == MODE_HARD)
            for (size_t i = 0; i < cs.size(); ++i) cs[i].ab += al[i];

        // candidate strains of the next level (lines 473-551)
        s.branching = false;
        struct Child { Cand c; int parent; };
        std::vector<Child> kids;
        for (size_t i = 0; i < cs.size(); ++i)
        {
            const int v = cs[i].node;
            const int e0 = g.out_off[v], e1 = g.out_off[v + 1];
            double oz = 0, moc = 0;
            for (int e2 = e0; e2 < e1; ++e2) { oz += g.out_cover[e2]; if (moc < g.out_cover[e2]) moc = g.out_cover[e2]; }
            int dd = 0;
            for (int e2 = e0; e2 < e1; ++e2)
            {
                const int o = g.out_to[e2];
                const double oc = g.out_cover[e2];
                Child k;
                k.parent = (int)i;
                k.c = cs[i];
                if (o != g.end_node && oz > 0)
                {
                    if (oc <= 1. && oc < moc) { dd += 1; continue; }
                    k.c.ab = (oc > 0) ? cs[i].ab * oc / oz : oz * std::min(0.01, tau);
                }
                k.c.node = o;
                k.c.tail = (int)s.trail.size();
                s.trail.push_back({cs[i].tail, o});
                k.c.hash = extend_hash(cs[i].hash, g.label_chars.data() + g.label_off[o], g.label_off[o + 1] - g.label_off[o]);
                k.c.len = cs[i].len + (uint64_t)(g.label_off[o + 1] - g.label_off[o]);
                kids.push_back(k);
            }
            if (e1 - e0 > 1 + dd) s.branching = true;
        }
        if (kids.size() > 80)
        {
            std::vector<double> ssa;
            for (const Child& k : kids) ssa.push_back(k.c.ab);
            std::sort(ssa.begin(), ssa.end(), [](double x, double y) { return x > y; });
            const double cut = ssa[80];
            std::vector<Child> keep;
            for (const Child& k : kids) if (!(k.c.ab < cut)) keep.push_back(k);
            kids.swap(keep);
        }
        // slots: the first surviving child of a parent takes the parent's slot, the others copy it
        std::vector<char> parent_used(cs.size(), 0);
        std::vector<Cand> next;
        next.reserve(kids.size());
        for (Child& k : kids)
        {
            if (!parent_used[k.parent]) { parent_used[k.parent] = 1; k.c.slot = cs[k.parent].slot; }
            else
            {
                const int dst = take_slot(s);
                inherit(s, cs[k.parent].slot, dst);
                k.c.slot = dst;
            }
            next.push_back(k.c);
        }
        for (size_t i = 0; i < cs.size(); ++i) if (!parent_used[i]) give_slot(s, cs[i].slot);
        cs.swap(next);
        s.cur.swap(s.nxt);
        s.epoch += 1;
        s.levels += 1;
        if (s.cur.empty()) s.done = true;
        else if (s.levels > g.n_nodes + 1)
        {   // the level walk of a DAG ends within n_nodes levels; anything longer is a malformed graph
            s.status = RAMBL_ERR_INVALID;
            s.have_result = false;
            s.failed = true;
            s.done = true;
        }
    }

    // pack the private shares of the subgroups into the pinned arenas (offsets become arena offsets);
    // the offsets are a serial prefix sum, the copies run on the workers
    void pack_step()
    {
        h_groups.clear();
        group_sub.clear();
        w_total = 0;
        size_t ti = 0, td = 0;
        std::vector<size_t> at_i, at_d;
        for (size_t i = 0; i < subs.size(); ++i)
        {
            Sub& s = subs[i];
            if (!s.has_group) continue;
            StepGroup sg = s.sg;
            const int bi = (int)ti, bd = (int)td;
            at_i.push_back(ti);
            at_d.push_back(td);
            ti += s.I.size();
            td += s.Dv.size();
            if (ti > 0x7fffffffull) throw Error(RAMBL_ERR_CAPACITY, "a level of this batch needs more than 2^31 staged integers");
            sg.slot_off += bi; sg.lab_off += bi; sg.rid_off += bi; sg.draw_off += bi;
            sg.ab_off += bd;
            sg.w_off = w_total;
            w_total += scratch_doubles(sg.S, sg.D);  // weights, normalisers (k_hard), letter codes (k_gibbs)
            s.ab_off = sg.ab_off;
            h_groups.push_back(sg);
            group_sub.push_back((int)i);
            stats.loglik_updates += s.step_updates;
            stats.draws += s.step_draws;
            s.has_group = false;
        }
        n_I = ti;
        n_D = td;
        // one pinned arena per step -- [doubles | step groups | ints], each part 16-byte aligned -- so that a
        // level costs one host-to-device copy
        off_groups = (sizeof(double) * td + 15) & ~size_t(15);
        off_I = (off_groups + sizeof(StepGroup) * h_groups.size() + 15) & ~size_t(15);
        arena_bytes = off_I + sizeof(int) * ti;
        p_arena.reserve(std::max<size_t>(arena_bytes, 16));
        int* const pI = reinterpret_cast<int*>(p_arena.p + off_I);
        double* const pD = reinterpret_cast<double*>(p_arena.p);
        auto copy_one = [&](size_t k) {
            const Sub& s = subs[group_sub[k]];
            if (!s.I.empty()) memcpy(pI + at_i[k], s.I.data(), sizeof(int) * s.I.size());
            if (!s.Dv.empty()) memcpy(pD + at_d[k], s.Dv.data(), sizeof(double) * s.Dv.size());
        };
        if (workers && group_sub.size() >= 8) workers->run(group_sub.size(), copy_one);
        else for (size_t k = 0; k < group_sub.size(); ++k) copy_one(k);
    }

    // Slot copies queued since the last flush.  The copies of one launch run in no defined order, so a copy that
    // reads or overwrites a slot an EARLIER queued copy writes or reads (a candidate taken at "$" right after it
    // was created as a second child, say) waits for the next launch: ops are dealt into waves by dependency.
    void flush_inherits()
    {
        std::vector<std::vector<InheritOp>> waves;
        std::vector<int> wave_of;
        for (size_t i = 0; i < subs.size(); ++i)
        {
            Sub& s = subs[i];
            wave_of.assign(s.ops.size(), 0);
            for (size_t a = 0; a < s.ops.size(); ++a)
            {
                int w = 0;
                for (size_t b = 0; b < a; ++b)
                {
                    const bool raw = s.ops[b].second == s.ops[a].first;   // reads what b writes
                    const bool waw = s.ops[b].second == s.ops[a].second;  // overwrites what b writes
                    const bool war = s.ops[b].first == s.ops[a].second;   // overwrites what b reads
                    if (raw || waw || war) w = std::max(w, wave_of[b] + 1);
                }
                wave_of[a] = w;
                if ((int)waves.size() <= w) waves.resize(w + 1);
                waves[w].push_back({s.ll.p, (long long)s.R, s.sub.p, s.ops[a].first, s.ops[a].second});
            }
            s.ops.clear();
            stats.launches += s.local_launches;
            s.local_launches = 0;
        }
        for (const std::vector<InheritOp>& ops : waves)
            for (size_t b = 0; b < ops.size(); b += 32768)
            {
                const int n = (int)std::min<size_t>(32768, ops.size() - b);
                d_ops.reserve(n);
                RAMBL_CUDA(cudaMemcpyAsync(d_ops.p, ops.data() + b, sizeof(InheritOp) * n, cudaMemcpyHostToDevice, st));
                stats.h2d_bytes += (long long)sizeof(InheritOp) * n;
                launch_inherit(d_ops.p, n, max_stride, st, &stats.launches);
                // the staging vector is pageable (the copy has left it when cudaMemcpyAsync returns) and a
                // second chunk reuses d_ops in stream order
            }
    }

    const double* run_step()
    {
        flush_inherits();
        if (h_groups.empty()) return nullptr;
        int max_S = 0, max_m = 0, max_D = 0;
        bool any_hard = false, any_gibbs = false;
        for (size_t k = 0; k < h_groups.size(); ++k)
        {
            StepGroup& sg = h_groups[k];
            Sub& s = subs[group_sub[k]];
            sg.ll = s.ll.p;
            sg.ll_stride = s.R;
            sg.sub = s.sub.p;
            sg.label_chars = s.d_label;
            sg.pool_chars = s.d_pool;
            max_S = std::max(max_S, sg.S);
            max_m = std::max(max_m, sg.m);
            max_D = std::max(max_D, sg.D);
            any_hard = any_hard || sg.mode == MODE_HARD;
            any_gibbs = any_gibbs || sg.mode == MODE_GIBBS || sg.mode == MODE_ASSIGN;
        }
        d_arena.reserve(std::max<size_t>(arena_bytes, 16));
        d_W.reserve(std::max<long long>(1, w_total));
        p_al.reserve(std::max<size_t>(1, n_D));
        memcpy(p_arena.p + off_groups, h_groups.data(), sizeof(StepGroup) * h_groups.size());
        RAMBL_CUDA(cudaMemcpyAsync(d_arena.p, p_arena.p, arena_bytes, cudaMemcpyHostToDevice, st));
        double* const dD = reinterpret_cast<double*>(d_arena.p);
        StepLaunch L;
        L.groups = reinterpret_cast<const StepGroup*>(d_arena.p + off_groups); L.n_groups = (int)h_groups.size();
        L.iarena = reinterpret_cast<const int*>(d_arena.p + off_I); L.darena = dD;
        L.weights = d_W.p; L.uniforms = d_U.p; L.n_uniforms = kUniforms; L.counters = d_counters.p;
        L.max_S = max_S; L.max_m = max_m; L.max_D = max_D; L.any_hard = any_hard; L.any_gibbs = any_gibbs;
        if (any_gibbs)
        {
            cudaEvent_t a, b;
            RAMBL_CUDA(cudaEventCreate(&a));
            RAMBL_CUDA(cudaEventCreate(&b));
            gibbs_events.push_back(a);
            gibbs_events.push_back(b);
            L.gibbs_begin = a;
            L.gibbs_end = b;
            stats.gibbs_launches += 1;
            for (const StepGroup& sg : h_groups)
                if ((sg.mode == MODE_GIBBS || sg.mode == MODE_ASSIGN) && sg.S >= 2)
                    stats.gibbs_bytes += (long long)sg.nsweeps * sg.D * (sg.S + 1) * 8;
        }
        stats.h2d_bytes += (long long)(sizeof(StepGroup) * h_groups.size() + sizeof(int) * n_I + sizeof(double) * n_D);
        stats.d2h_bytes += (long long)(sizeof(double) * n_D);
        launch_level_step(L, st, &stats.launches);
        RAMBL_CUDA(cudaMemcpyAsync(p_al.p, dD, sizeof(double) * n_D, cudaMemcpyDeviceToHost, st));
        RAMBL_CUDA(cudaStreamSynchronize(st));
        stats.level_steps += 1;
        return p_al.p;
    }

    void collect_kernel_times()
    {
        for (size_t k = 0; k + 1 < gibbs_events.size(); k += 2)
        {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, gibbs_events[k], gibbs_events[k + 1]) == cudaSuccess) stats.gibbs_ms += ms;
            cudaEventDestroy(gibbs_events[k]);
            cudaEventDestroy(gibbs_events[k + 1]);
        }
        gibbs_events.clear();
        unsigned long long c[4] = {0, 0, 0, 0};
        if (d_counters.p && cudaMemcpy(c, d_counters.p, sizeof c, cudaMemcpyDeviceToHost) == cudaSuccess)
        {
            stats.gibbs_rounds += (long long)c[0];
            stats.gibbs_passes += (long long)c[1];
            if (c[2]) throw Error(RAMBL_ERR_CUDA, "internal error: a Gibbs round did not settle (non-finite weights?)");
        }
    }


    // ---- read_assign for every subgroup in one step, NonparametricClustering.cpp:776-836 ----
    const double* assign_step()
    {
        for (size_t i = 0; i < subs.size(); ++i)
        {
            Sub& s = subs[i];
            s.mode = MODE_NONE;
            s.has_group = false;
            s.I.clear();
            s.Dv.clear();
            s.step_updates = s.step_draws = 0;
            std::vector<int>& h_I = s.I;
            std::vector<double>& h_D = s.Dv;
            if (s.status != RAMBL_OK || s.result.empty()) continue;
            const SubgroupInput& in = *s.in;
            const int S = (int)s.result.size(), R = s.g->n_reads;
            int D = 0;
            for (int r = 0; r < R; ++r) D += in.read_cn[r];
            if (D == 0) continue;
            StepGroup sg;
            memset(&sg, 0, sizeof sg);
            sg.S = S; sg.m = R; sg.D = D; sg.mode = MODE_ASSIGN; sg.read_size = D;
            sg.nsweeps = std::min(prm.n, 40000 / D);
            sg.slot_off = (int)h_I.size();
            for (const Cand& c : s.result) h_I.push_back(c.slot);
            sg.lab_off = (int)h_I.size();
            h_I.resize(h_I.size() + 2 * (size_t)S, 0);
            sg.rid_off = (int)h_I.size();
            for (int r = 0; r < R; ++r) h_I.push_back(r);
            h_I.resize(h_I.size() + 3 * (size_t)R, 0);
            sg.draw_off = (int)h_I.size();
            h_I.resize(h_I.size() + 2 * (size_t)D);
            int* dr = &h_I[sg.draw_off];
            int* dm = dr + D;
            int d = 0;
            for (int r = 0; r < R; ++r)
                for (int cn = in.read_cn[r]; cn > 0; --cn, ++d)
                {
                    dr[d] = r;
                    dm[d] = in.pair_val[in.pair_off[r] + cn - 1];
                    if (dm[d] >= s.R) throw Error(RAMBL_ERR_INVALID, "mate id out of range");
                }
            sg.ab_off = (int)h_D.size();
            for (const Cand& c : s.result) h_D.push_back(c.ab);
            s.mode = MODE_ASSIGN;
            s.sg = sg;
            s.has_group = true;
            if (S >= 2) { s.step_draws = (long long)D * sg.nsweeps; s.draws += s.step_draws; }
        }
        pack_step();
        return run_step();
    }

    // ---- the device-resident walk (walk.cu): level tables, one launch, results -------------------------------
    struct WalkPlan  // host side of one subgroup's tables
    {
        bool eligible = false;
        bool handoff = false;            // "$" shares a level with other nodes: the walk stops before that level and the
        std::vector<int> handoff_cur;    // level-synchronous path goes on from there (these are the level's nodes)
        int reason = 0;  // why not: 1 a node on two levels, 2 "$" not alone / not last, 3 a read twice on a level, 4 entry range,
                         // 5 "^" carries reads, 6 size, 7 mate id out of range, 8 no graph
        int n_levels = 0, max_m = 0, max_D = 0;
        std::vector<int> lvl_moff;           // [n_levels] -1, or the level's start in the multi-letter tables
        long long n_ment = 0, n_mchars = 0;  // entries / letters of the levels that hold multi-letter entries
        int mixed_levels = 0;                // levels with a multi-letter read string next to a one-letter node
        std::vector<int> order;              // nodes in walk order (levels concatenated)
        std::vector<int> lvl_ent_off;        // [n_levels + 1]
        std::vector<unsigned char> lvl_dup;  // [n_levels] a read with several entries on the level
        long long n_ent = 0, n_chars = 0;
        // offsets into the static arena (bytes) and the scratch arena (bytes)
        size_t o_label_off = 0, o_out_off = 0, o_out_to = 0, o_out_cover = 0, o_lvl = 0, o_dup = 0, o_rid = 0, o_cn = 0, o_char1 = 0, o_moff = 0,
               o_soff = 0, o_len = 0, o_chars = 0, o_pair_off = 0, o_pair_val = 0;
        size_t d_present = 0, d_free = 0, d_cand0 = 0, d_cand1 = 0, d_trail = 0, d_W = 0, d_doff = 0, d_dent = 0, d_dmate = 0,
               d_fresh = 0, d_ab = 0, d_ops = 0, d_kid = 0, d_lut = 0, d_helper = 0, d_res = 0, d_paths = 0, d_fslot = 0, d_fab = 0;
        int trail_cap = 0;
    };
    std::vector<WalkPlan> plans;
    PinBuf<char> p_static;
    DevBuf<char> d_static, d_scratch;
    DevBuf<WalkSub> d_walk;
    DevBuf<WalkResult> d_walk_res;
    // Subgroups the walk cannot take are solved by the level-synchronous path WHILE the walk kernel runs: a second call
    // of infer_batch on its own stream and host thread (their per-level launches slip into the SMs the walk's waves
    // leave free, instead of adding a whole chain's latency after it)
    std::vector<SubgroupInput> side_in;
    std::vector<SubgroupResult> side_out;
    EngineStats side_stats;
    std::thread side_thread;
    std::exception_ptr side_error;
    cudaStream_t side_stream = nullptr;
    void join_side()
    {
        if (side_thread.joinable()) side_thread.join();
        if (side_stream) { cudaStreamDestroy(side_stream); side_stream = nullptr; }
        if (side_error) { std::exception_ptr e2 = side_error; side_error = nullptr; std::rethrow_exception(e2); }
    }
    ~Engine() { if (side_thread.joinable()) side_thread.join(); if (side_stream) cudaStreamDestroy(side_stream); }
    DevBuf<int> d_final_slot;
    DevBuf<double> d_final_ab;

    // The walk order of prepare() -- cur = {0}; next = the successors of cur not yet listed for the next level, in
    // order -- is a property of the graph, not of the candidate strains.  The device walk takes a subgroup when "$" ends
    // the walk alone on its level and the read-pool entries fit the compact tables; a node reached on two levels
    // simply has its pool listed on both (as prepare() does), a read with several entries on one level has the later
    // ones flagged (they are added after the first, in entry order).  Anything else goes through the
    // level-synchronous path.
    void plan_walk(size_t i) { plan_walk_graph(*subs[i].g, *subs[i].in, subs[i].R, plans[i]); }

    // (a function of the graph and the read pairs alone: walk_eligibility() below answers without a device)
    static void plan_walk_graph(const FlatGraph& g, const SubgroupInput& in, int R, WalkPlan& p)
    {
        p = WalkPlan();
        if (g.n_nodes < 2 || g.end_node < 0) { p.reason = 8; return; }
        std::vector<int> mark(g.n_nodes, -1);
        std::vector<int> cur(1, 0), nxt;
        std::vector<int> seen_rid(std::max(1, g.n_reads), -1), mult(std::max(1, g.n_reads), 0);
        p.lvl_ent_off.assign(1, 0);
        int level = 0;
        bool ok = true, ended = false;
        while (!cur.empty() && ok)
        {
            if (ended) { ok = false; p.reason = 2; break; }  // something follows "$"
            if (level > g.n_nodes + 1) { ok = false; p.reason = 1; break; }  // not a DAG
            if (cur.size() > 1 && std::find(cur.begin(), cur.end(), g.end_node) != cur.end())
            {   // "$" next to other nodes (a graph that is not strictly levelled): the device walks up to here
                p.handoff = true;
                p.handoff_cur = cur;
                ended = true;
                level += 1;
                break;
            }
            long long m = 0, D = 0, chars = 0;
            unsigned char dup = 0;
            bool multi = false, single_node = false;
            nxt.clear();
            for (int u : cur)
            {
                if (u == g.end_node) ended = true;
                else if (u != 0)
                {
                    if (g.label_off[u + 1] - g.label_off[u] == 1) single_node = true;
                    for (int e = g.pool_off[u]; e < g.pool_off[u + 1]; ++e)
                    {
                        const int rid = g.pool_rid[e], cn = g.pool_cn[e], len = g.pool_str_off[e + 1] - g.pool_str_off[e];
                        if (cn < 1 || cn > 255 || len < 1 || len > 255) { ok = false; p.reason = 4; break; }
                        // further entries of a read on one level are added after the first, in entry order
                        if (seen_rid[rid] == level) { mult[rid] += 1; dup = (unsigned char)std::min(255, std::max<int>(dup, mult[rid])); }
                        else { seen_rid[rid] = level; mult[rid] = 0; }
                        if (len > 1) multi = true;
                        m += 1; D += cn; chars += len;
                    }
                }
                else if (g.pool_off[1] != g.pool_off[0] || level != 0) { ok = false; p.reason = 5; }  // "^" carries no reads
                if (!ok) break;
                p.order.push_back(u);
                for (int e = g.out_off[u]; e < g.out_off[u + 1]; ++e)
                {
                    const int v = g.out_to[e];
                    if (mark[v] != level) { mark[v] = level; nxt.push_back(v); }
                }
            }
            if (!ok) break;
            p.n_ent += m;
            p.lvl_ent_off.push_back((int)p.n_ent);
            p.lvl_dup.push_back(dup);
            p.lvl_moff.push_back(multi ? (int)p.n_ment : -1);
            if (multi) { p.n_ment += m; p.n_mchars += chars; }
            if (multi && single_node) p.mixed_levels += 1;
            p.max_m = std::max<long long>(p.max_m, m);
            p.max_D = std::max<long long>(p.max_D, D);
            if (D > 40000 * 8 || p.n_ent > 0x7fffff00LL) { ok = false; p.reason = 6; break; }
            cur.swap(nxt);
            level += 1;
        }
        if (!ok || !ended || level < 2) { if (!p.reason) p.reason = 2; return; }
        for (int v : in.pair_val) if (v >= R) { p.reason = 7; return; }  // reported by the level-synchronous path
        p.n_levels = level;
        p.eligible = true;
    }

    static size_t up16(size_t b) { return (b + 15) & ~size_t(15); }

    // Returns the number of subgroups handed to the device walk; on return their Sub holds the closed result.
    int run_walk(int forced_nb, int forced_tile, int forced_cluster)
    {
        const size_t n = subs.size();
        const auto w0 = std::chrono::steady_clock::now();
        auto lap = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count(); };
        double ms_plan = 0, ms_fill = 0, ms_launch = 0, ms_kernel_done = 0;
        plans.assign(n, WalkPlan());
        auto each = [&](const std::function<void(size_t)>& fn) {
            if (workers) workers->run(n, fn);
            else for (size_t i = 0; i < n; ++i) fn(i);
        };
        each([&](size_t i) { plan_walk(i); });
        ms_plan = lap();
        for (const WalkPlan& p : plans) stats.offtable_levels += p.mixed_levels;
        std::vector<int> take;
        size_t stat_bytes = 0, scr_bytes = 0;
        for (size_t i = 0; i < n; ++i)
        {
            WalkPlan& p = plans[i];
            if (!p.eligible) continue;
            const Sub& s = subs[i];
            const FlatGraph& g = *s.g;
            const SubgroupInput& in = *s.in;
            take.push_back((int)i);
            auto put = [&](size_t& off, size_t bytes) { off = stat_bytes; stat_bytes += up16(bytes); };
            put(p.o_label_off, sizeof(int) * (g.n_nodes + 1));
            put(p.o_out_off, sizeof(int) * (g.n_nodes + 1));
            put(p.o_out_to, sizeof(int) * g.out_to.size());
            put(p.o_out_cover, sizeof(int) * g.out_to.size());
            put(p.o_lvl, sizeof(int) * p.lvl_ent_off.size());
            put(p.o_dup, p.lvl_dup.size());
            put(p.o_rid, sizeof(unsigned) * p.n_ent);
            put(p.o_cn, (size_t)p.n_ent);
            put(p.o_char1, (size_t)p.n_ent);
            put(p.o_moff, sizeof(int) * p.lvl_moff.size());
            put(p.o_soff, sizeof(unsigned) * p.n_ment);
            put(p.o_len, (size_t)p.n_ment);
            put(p.o_chars, (size_t)p.n_mchars);
            put(p.o_pair_off, sizeof(int) * in.pair_off.size());
            put(p.o_pair_val, sizeof(int) * in.pair_val.size());
            auto scr = [&](size_t& off, size_t bytes) { off = scr_bytes; scr_bytes += up16(bytes) + 112; scr_bytes &= ~size_t(127); };
            p.trail_cap = p.n_levels * 160 + 2048;
            const size_t Dp = ((size_t)p.max_D + 31) & ~size_t(31);
            scr(p.d_present, (size_t)s.R);
            scr(p.d_free, sizeof(int) * s.slot_cap);
            scr(p.d_cand0, sizeof(WalkCand) * WALK_KMAX);
            scr(p.d_cand1, sizeof(WalkCand) * WALK_KMAX);
            scr(p.d_trail, sizeof(int2) * (size_t)p.trail_cap);
            scr(p.d_W, sizeof(double) * ((size_t)WALK_SMAX * Dp + 2 * (size_t)p.max_D + 64));
            scr(p.d_doff, sizeof(int) * ((size_t)p.max_m + 1));
            scr(p.d_dent, sizeof(int) * (size_t)std::max(1, p.max_D));
            scr(p.d_dmate, sizeof(int) * (size_t)std::max(1, p.max_D));
            scr(p.d_fresh, (size_t)std::max(1, p.max_m));
            scr(p.d_ab, sizeof(double) * WALK_SMAX);
            scr(p.d_ops, sizeof(int2) * WALK_KMAX);
            scr(p.d_kid, sizeof(double) * WALK_KMAX);
            scr(p.d_lut, sizeof(double) * WALK_SMAX * 36);
            scr(p.d_helper, sizeof(int) * 8);
            scr(p.d_paths, sizeof(int) * (size_t)WALK_SMAX * p.n_levels);
        }
        if (getenv("RAMBL_TRACE") && take.size() < n)
        {
            int why[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (const WalkPlan& p : plans) if (!p.eligible) why[std::min(std::max(p.reason, 0), 8)] += 1;
            fprintf(stderr, "[rambl] device walk: %zu of %zu subgroups not eligible (reasons 1..8: %d %d %d %d %d %d %d %d)\n",
                    n - take.size(), n, why[1], why[2], why[3], why[4], why[5], why[6], why[7], why[8]);
            int shown = 0;
            for (size_t i = 0; i < n && shown < 4; ++i)
                if (!plans[i].eligible)
                {
                    fprintf(stderr, "[rambl]   subgroup %zu: reason %d after %zu levels, %d nodes, %d reads\n", i, plans[i].reason,
                            plans[i].lvl_ent_off.size() - 1, subs[i].g->n_nodes, subs[i].g->n_reads);
                    ++shown;
                }
        }
        if (take.empty()) return 0;
        // CTAs are handed to the SMs in launch order: the subgroups with the most graph nodes (more variant nodes, more
        // candidate strains, longer chains) go first, so that the last wave is made of the short walks
        if (!getenv("RAMBL_WALK_NOSORT"))
            std::stable_sort(take.begin(), take.end(), [&](int x, int y) { return subs[x].g->n_nodes > subs[y].g->n_nodes; });
        if (take.size() < n)
        {   // the rest: level-synchronous, concurrently (see side_in)
            for (size_t i = 0; i < n; ++i)
                if (!plans[i].eligible)
                {
                    subs[i].side = (int)side_in.size();
                    subs[i].done = true;
                    side_in.push_back(*subs[i].in);
                }
            int device = 0;
            RAMBL_CUDA(cudaGetDevice(&device));
            RAMBL_CUDA(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
            side_thread = std::thread([this, device] {
                try
                {
                    RAMBL_CUDA(cudaSetDevice(device));
                    InferParams p2 = prm;
                    p2.level_synchronous = true;
                    p2.on_device_phase = nullptr;
                    infer_batch(side_in, p2, side_out, side_stats, side_stream);
                }
                catch (...) { side_error = std::current_exception(); }
            });
        }
        // ---- fill the static tables (pinned) on the workers, one copy to the device
        p_static.reserve(stat_bytes);
        d_static.reserve(stat_bytes);
        d_scratch.reserve(scr_bytes);
        d_walk_res.reserve(take.size());
        d_final_slot.reserve(take.size() * WALK_SMAX);
        d_final_ab.reserve(take.size() * WALK_SMAX);
        char* const H = p_static.p;
        auto fill = [&](size_t k) {
            const size_t i = (size_t)take[k];
            const WalkPlan& p = plans[i];
            const Sub& s = subs[i];
            const FlatGraph& g = *s.g;
            const SubgroupInput& in = *s.in;
            memcpy(H + p.o_label_off, g.label_off.data(), sizeof(int) * (g.n_nodes + 1));
            memcpy(H + p.o_out_off, g.out_off.data(), sizeof(int) * (g.n_nodes + 1));
            if (!g.out_to.empty())
            {
                memcpy(H + p.o_out_to, g.out_to.data(), sizeof(int) * g.out_to.size());
                memcpy(H + p.o_out_cover, g.out_cover.data(), sizeof(int) * g.out_to.size());
            }
            memcpy(H + p.o_lvl, p.lvl_ent_off.data(), sizeof(int) * p.lvl_ent_off.size());
            memcpy(H + p.o_dup, p.lvl_dup.data(), p.lvl_dup.size());
            unsigned* rid = reinterpret_cast<unsigned*>(H + p.o_rid);
            unsigned char* cn = reinterpret_cast<unsigned char*>(H + p.o_cn);
            char* char1 = H + p.o_char1;
            unsigned* soff = reinterpret_cast<unsigned*>(H + p.o_soff);
            unsigned char* len = reinterpret_cast<unsigned char*>(H + p.o_len);
            char* chars = H + p.o_chars;
            memcpy(H + p.o_moff, p.lvl_moff.data(), sizeof(int) * p.lvl_moff.size());
            size_t at = 0, at_m = 0, at_c = 0;
            std::vector<int> seen(std::max(1, g.n_reads), -1);
            size_t lvl = 0;
            for (int u : p.order)
            {
                if (u == 0 || u == g.end_node) continue;
                for (int e = g.pool_off[u]; e < g.pool_off[u + 1]; ++e, ++at)
                {
                    while (lvl + 1 < p.lvl_ent_off.size() && (long long)at >= p.lvl_ent_off[lvl + 1]) ++lvl;
                    const int r0 = g.pool_rid[e];
                    rid[at] = (unsigned)r0 | (seen[r0] == (int)lvl ? 0x80000000u : 0u);
                    seen[r0] = (int)lvl;
                    cn[at] = (unsigned char)g.pool_cn[e];
                    const int l = g.pool_str_off[e + 1] - g.pool_str_off[e];
                    char1[at] = g.pool_chars[g.pool_str_off[e]];
                    if (p.lvl_moff[lvl] >= 0)
                    {
                        soff[at_m] = (unsigned)at_c;
                        len[at_m] = (unsigned char)l;
                        memcpy(chars + at_c, g.pool_chars.data() + g.pool_str_off[e], (size_t)l);
                        at_c += (size_t)l;
                        ++at_m;
                    }
                }
            }
            memcpy(H + p.o_pair_off, in.pair_off.data(), sizeof(int) * in.pair_off.size());
            if (!in.pair_val.empty()) memcpy(H + p.o_pair_val, in.pair_val.data(), sizeof(int) * in.pair_val.size());
        };
        if (workers && take.size() >= 4) workers->run(take.size(), fill);
        else for (size_t k = 0; k < take.size(); ++k) fill(k);
        ms_fill = lap();
        RAMBL_CUDA(cudaMemcpyAsync(d_static.p, H, stat_bytes, cudaMemcpyHostToDevice, st));
        stats.h2d_bytes += (long long)stat_bytes;
        // ---- descriptors
        std::vector<WalkSub> hs(take.size());
        char* const Dst = d_static.p;
        char* const Dsc = d_scratch.p;
        int max_levels = 0;
        for (size_t k = 0; k < take.size(); ++k)
        {
            const size_t i = (size_t)take[k];
            const WalkPlan& p = plans[i];
            const Sub& s = subs[i];
            WalkSub& w = hs[k];
            memset(&w, 0, sizeof w);
            w.label_off = reinterpret_cast<const int*>(Dst + p.o_label_off);
            w.label_chars = s.d_label;
            w.out_off = reinterpret_cast<const int*>(Dst + p.o_out_off);
            w.out_to = reinterpret_cast<const int*>(Dst + p.o_out_to);
            w.out_cover = reinterpret_cast<const int*>(Dst + p.o_out_cover);
            w.end_node = s.g->end_node;
            w.n_levels = p.n_levels;
            w.lvl_ent_off = reinterpret_cast<const int*>(Dst + p.o_lvl);
            w.lvl_dup = reinterpret_cast<const unsigned char*>(Dst + p.o_dup);
            w.ent_rid = reinterpret_cast<const unsigned*>(Dst + p.o_rid);
            w.ent_cn = reinterpret_cast<const unsigned char*>(Dst + p.o_cn);
            w.ent_char1 = Dst + p.o_char1;
            w.lvl_moff = reinterpret_cast<const int*>(Dst + p.o_moff);
            w.m_soff = reinterpret_cast<const unsigned*>(Dst + p.o_soff);
            w.m_len = reinterpret_cast<const unsigned char*>(Dst + p.o_len);
            w.m_chars = Dst + p.o_chars;
            w.pair_off = reinterpret_cast<const int*>(Dst + p.o_pair_off);
            w.pair_val = reinterpret_cast<const int*>(Dst + p.o_pair_val);
            w.R = s.R;
            w.slot_cap = s.slot_cap;
            w.ll = s.ll.p;
            w.sub = s.sub.p;
            w.present = reinterpret_cast<unsigned char*>(Dsc + p.d_present);
            w.free_slots = reinterpret_cast<int*>(Dsc + p.d_free);
            w.cand[0] = reinterpret_cast<WalkCand*>(Dsc + p.d_cand0);
            w.cand[1] = reinterpret_cast<WalkCand*>(Dsc + p.d_cand1);
            w.trail = reinterpret_cast<int2*>(Dsc + p.d_trail);
            w.trail_cap = p.trail_cap;
            w.W = reinterpret_cast<double*>(Dsc + p.d_W);
            w.ent_doff = reinterpret_cast<int*>(Dsc + p.d_doff);
            w.draw_entry = reinterpret_cast<int*>(Dsc + p.d_dent);
            w.draw_mate = reinterpret_cast<int*>(Dsc + p.d_dmate);
            w.fresh = reinterpret_cast<unsigned char*>(Dsc + p.d_fresh);
            w.ab_io = reinterpret_cast<double*>(Dsc + p.d_ab);
            w.ops = reinterpret_cast<int2*>(Dsc + p.d_ops);
            w.kid_ab = reinterpret_cast<double*>(Dsc + p.d_kid);
            w.lut = reinterpret_cast<double*>(Dsc + p.d_lut);
            w.helper = reinterpret_cast<int*>(Dsc + p.d_helper);
            w.res = d_walk_res.p + k;
            w.paths = reinterpret_cast<int*>(Dsc + p.d_paths);
            w.final_slot = d_final_slot.p + k * WALK_SMAX;
            w.final_ab = d_final_ab.p + k * WALK_SMAX;
            RAMBL_CUDA(cudaMemsetAsync(w.present, 0, (size_t)s.R, st));
            max_levels = std::max(max_levels, p.n_levels);
        }
        d_walk.reserve(hs.size());
        RAMBL_CUDA(cudaMemcpyAsync(d_walk.p, hs.data(), sizeof(WalkSub) * hs.size(), cudaMemcpyHostToDevice, st));
        stats.h2d_bytes += (long long)(sizeof(WalkSub) * hs.size());
        // ---- CTA shape.  Measured on 500 subgroups: eight warps per subgroup with one subgroup per SM
        // beat two or four warps with all subgroups resident at once -- a chain is latency-bound, and a round that
        // speculates over 256 draws makes up for the waves -- so the CTA is always eight warps wide, with the widest
        // weight tiles that fit next to the per-strain arrays.
        // A batch with fewer subgroups than SMs gives every subgroup a CLUSTER of CTAs: the extra CTAs join the Gibbs
        // chains (more 32-draw blocks per round), the one sequential part of a level.
        const size_t sm_bytes = 227 * 1024;
        int nb = forced_nb > 0 ? forced_nb : 8, tile = 8;
        int cluster = 1;
        if (nb == 8)
        {
            // (measured on configs[1]: 8 CTAs 2.2x, 4 CTAs 1.6x, 2 CTAs 1.05x the single CTA -- the per-pass exchange and
            // the cluster barriers eat most of what two CTAs gain)
            if (take.size() <= 16) cluster = 8;
            else if (take.size() <= 32) cluster = 4;
            else if (take.size() <= 72) cluster = 2;
            if (prm.max_cluster > 0) cluster = std::min(cluster, prm.max_cluster);
            if (forced_cluster > 0) cluster = forced_cluster;
        }
        while (tile + 4 <= WALK_SMAX && walk_smem_bytes(nb, tile + 4, cluster > 1) <= sm_bytes) tile += 4;
        if (forced_tile > 0) tile = forced_tile;
        WalkParams wp;
        wp.n = prm.n;
        wp.single_buffer = getenv("RAMBL_WALK_TILES") ? atoi(getenv("RAMBL_WALK_TILES")) : 3;
        wp.tau = tau;
        wp.uniforms = d_U.p;
        wp.counters = d_counters.p;
        cudaEvent_t e0, e1;
        RAMBL_CUDA(cudaEventCreate(&e0));
        RAMBL_CUDA(cudaEventCreate(&e1));
        RAMBL_CUDA(cudaEventRecord(e0, st));
        launch_walk(d_walk.p, (int)take.size(), wp, nb, tile, cluster, st, &stats.launches);
        RAMBL_CUDA(cudaEventRecord(e1, st));
        ms_launch = lap();
        if (prm.on_device_phase) prm.on_device_phase();
        // ---- results.  Wait for the kernel FIRST: an asynchronous copy into pageable host memory queued behind a running
        // kernel keeps the calling thread inside the driver until the kernel ends, and (measured) the CUDA calls of
        // every other thread of the process wait with it -- the next chunk of an overlapped solve could not set up its
        // own walk while this one ran.  cudaStreamSynchronize holds nothing.
        RAMBL_CUDA(cudaStreamSynchronize(st));
        std::vector<WalkResult> res(take.size());
        std::vector<int> all_slot(take.size() * WALK_SMAX);
        std::vector<double> all_ab(take.size() * WALK_SMAX);
        RAMBL_CUDA(cudaMemcpyAsync(res.data(), d_walk_res.p, sizeof(WalkResult) * res.size(), cudaMemcpyDeviceToHost, st));
        RAMBL_CUDA(cudaMemcpyAsync(all_slot.data(), d_final_slot.p, sizeof(int) * all_slot.size(), cudaMemcpyDeviceToHost, st));
        RAMBL_CUDA(cudaMemcpyAsync(all_ab.data(), d_final_ab.p, sizeof(double) * all_ab.size(), cudaMemcpyDeviceToHost, st));
        RAMBL_CUDA(cudaStreamSynchronize(st));
        ms_kernel_done = lap();
        stats.d2h_bytes += (long long)(sizeof(WalkResult) * res.size() + 12 * all_slot.size());
        float ms = 0;
        RAMBL_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        stats.walk_ms += ms;
        stats.walk_launches += 1;
        stats.level_steps += max_levels;
        if (getenv("RAMBL_TRACE"))
            fprintf(stderr, "[rambl] device walk: %zu subgroups, %d warps per CTA, %d CTAs per subgroup, tiles of %d strains, %zu B shared, %.1f ms, "
                            "tables %.1f MB, scratch %.1f MB\n", take.size(), nb, cluster, tile, walk_smem_bytes(nb, tile, cluster > 1), ms,
                    stat_bytes / 1e6, scr_bytes / 1e6);
        std::vector<std::vector<int>> h_paths(take.size());
        for (size_t k = 0; k < take.size(); ++k)
        {
            const WalkResult& r = res[k];
            if (r.status == WALK_NOT_SETTLED) throw Error(RAMBL_ERR_CUDA, "internal error: a Gibbs round did not settle (non-finite weights?)");
            if (r.status != WALK_DONE || r.n_cands <= 0) continue;
            const int nl = plans[take[k]].n_levels;
            h_paths[k].resize((size_t)r.n_cands * nl);
            RAMBL_CUDA(cudaMemcpyAsync(h_paths[k].data(), hs[k].paths, sizeof(int) * h_paths[k].size(), cudaMemcpyDeviceToHost, st));
            stats.d2h_bytes += (long long)(sizeof(int) * h_paths[k].size());
        }
        RAMBL_CUDA(cudaStreamSynchronize(st));
        if (getenv("RAMBL_TRACE"))
        {
            long long sum_S = 0, glev = 0, unst = 0;
            int max_S = 0, hist[WALK_NOT_RUN + 1] = {0};
            for (const WalkResult& r : res)
            {
                sum_S += r.sum_S; glev += r.gibbs_levels; unst += r.unstaged_levels; max_S = std::max(max_S, r.max_S);
                hist[std::min(std::max(r.status, 0), (int)WALK_NOT_RUN)] += 1;
            }
            fprintf(stderr, "[rambl] device walk: %lld Gibbs levels, mean %.1f strains, max %d, %lld levels unstaged; status counts:",
                    glev, glev ? (double)sum_S / glev : 0.0, max_S, unst);
            for (int k = 0; k <= WALK_NOT_RUN; ++k) fprintf(stderr, " %d", hist[k]);
            fprintf(stderr, "\n");
        }
        int taken = 0;
        std::vector<size_t> redo;
        std::vector<char> solved(take.size(), 0);
        for (size_t k = 0; k < take.size(); ++k)
        {
            const WalkResult& r = res[k];
            if (r.status >= WALK_TOO_MANY_STRAINS) { redo.push_back(take[k]); continue; }
            stats.draws += r.draws;
            stats.loglik_updates += r.loglik_updates;
            stats.walk_bytes += r.gibbs_bytes + 16 * r.loglik_updates + 16 * r.weight_pairs;
            stats.gibbs_bytes += r.gibbs_bytes;
            solved[k] = 1;
        }
        // the candidates and their paths, and "$" (sort + merge_strains: the strains keep their slots, the walk is over):
        // per subgroup, nothing shared -- on the workers
        auto close_one = [&](size_t k) {
            if (!solved[k]) return;
            const WalkResult& r = res[k];
            Sub& s = subs[take[k]];
            s.draws += r.draws;
            s.levels = plans[take[k]].n_levels;
            s.done = true;
            s.cands.clear();
            s.trail.clear();
            const int nl = plans[take[k]].n_levels;
            s.trail.reserve((size_t)std::max(r.n_cands, 0) * (size_t)nl);
            for (int c = 0; c < r.n_cands; ++c)
            {
                Cand cd;
                cd.slot = all_slot[k * WALK_SMAX + c];
                cd.ab = all_ab[k * WALK_SMAX + c];
                int prev = -1;
                for (int q = 0; q < nl; ++q) { s.trail.push_back({prev, h_paths[k][(size_t)c * nl + q]}); prev = (int)s.trail.size() - 1; }
                cd.tail = prev;
                cd.node = s.g->end_node;
                s.cands.push_back(cd);
            }
            if (!plans[take[k]].handoff) close_result(s, false);
        };
        if (workers && take.size() >= 4) workers->run(take.size(), close_one);
        else for (size_t k = 0; k < take.size(); ++k) close_one(k);
        for (size_t k = 0; k < take.size(); ++k)
        {
            if (!solved[k]) continue;
            if (!plans[take[k]].handoff) { ++taken; continue; }
            const WalkResult& r = res[k];
            Sub& s = subs[take[k]];
            const int nl = plans[take[k]].n_levels;
            // the walk stopped in front of a level that holds "$" next to other nodes: the level-synchronous path takes the
            // subgroup over exactly there (candidates, presence flags, free slots, the level's nodes)
            {
                const WalkSub& w = hs[k];
                std::vector<WalkCand> wc(r.n_cands);
                std::vector<unsigned char> pres((size_t)s.R);
                std::vector<int> fs((size_t)std::max(r.free_top, 0));
                if (r.n_cands) RAMBL_CUDA(cudaMemcpy(wc.data(), w.cand[r.cand_buf & 1], sizeof(WalkCand) * wc.size(), cudaMemcpyDeviceToHost));
                RAMBL_CUDA(cudaMemcpy(pres.data(), w.present, pres.size(), cudaMemcpyDeviceToHost));
                if (!fs.empty()) RAMBL_CUDA(cudaMemcpy(fs.data(), w.free_slots, sizeof(int) * fs.size(), cudaMemcpyDeviceToHost));
                for (int c = 0; c < r.n_cands; ++c)
                {
                    s.cands[c].node = wc[c].node;
                    s.cands[c].hash = wc[c].hash;
                    s.cands[c].len = wc[c].len;
                }
                s.present.assign(pres.begin(), pres.end());
                s.free_slots.assign(fs.begin(), fs.end());
                s.cur = plans[take[k]].handoff_cur;
                s.mark.assign(s.g->n_nodes, -1);
                s.epoch = nl;
                s.levels = nl - 1;
                s.branching = r.branching != 0;
                s.done = s.cands.empty();
                if (s.done) close_result(s, false);
            }
        }
        // subgroups the kernel gave up on start over on the level-synchronous path: wipe what the walk wrote
        for (size_t i : redo)
        {
            Sub& s = subs[i];
            RAMBL_CUDA(cudaMemsetAsync(s.ll.p, 0, sizeof(double) * (size_t)s.slot_cap * s.R, st));
            launch_init_models(s.sub.p, s.slot_cap, e, st, &stats.launches);
        }
        if (!redo.empty() && getenv("RAMBL_TRACE"))
            fprintf(stderr, "[rambl] device walk: %zu subgroups handed back to the level-synchronous path\n", redo.size());
        if (getenv("RAMBL_TRACE"))
            fprintf(stderr, "[rambl] device walk host ms: plans until %.1f, tables filled %.1f, launched %.1f, kernel and results back %.1f, closed %.1f\n",
                    ms_plan, ms_fill, ms_launch, ms_kernel_done, lap());
        return taken;
    }
};

}  // namespace

WalkEligibility walk_eligibility(const FlatGraph& g, const SubgroupInput& in)
{
    Engine::WalkPlan p;
    Engine::plan_walk_graph(g, in, std::max(1, g.n_reads), p);
    WalkEligibility e;
    e.eligible = p.eligible;
    e.handoff = p.handoff;
    e.reason = p.reason;
    e.levels = p.eligible ? p.n_levels : (int)p.lvl_ent_off.size() - 1;
    e.entries = p.n_ent;
    e.max_entries = p.max_m;
    e.max_draws = p.max_D;
    e.offtable_levels = p.mixed_levels;
    return e;
}

static int g_walk_mode = 1, g_walk_blocks = 0, g_walk_cluster = 0;
void set_walk_cluster(int c) { g_walk_cluster = c; }
int walk_cluster() { return g_walk_cluster; }
void set_walk_mode(int mode) { g_walk_mode = mode; }
int walk_mode() { return g_walk_mode; }
void set_walk_blocks(int nb) { g_walk_blocks = nb; }
int walk_blocks() { return g_walk_blocks; }

std::string strain_sequence(const FlatGraph& g, const std::vector<int>& path)
{
    std::string r;
    for (int u : path) r.append(g.label_chars.data() + g.label_off[u], g.label_chars.data() + g.label_off[u + 1]);
    return r;
}

std::string strain_plain_sequence(const FlatGraph& g, const std::vector<int>& path)
{
    std::string r;
    r.reserve(path.size());
    for (int u : path)
    {
        const char* l = g.label_chars.data() + g.label_off[u];
        const int n = g.label_off[u + 1] - g.label_off[u];
        if (n == 1 && (l[0] == '^' || l[0] == '$' || l[0] == '-' || l[0] == '=')) continue;  // the four one-letter marks
        r.append(l, (size_t)n);
    }
    return r;
}

void infer_batch(const std::vector<SubgroupInput>& in, const InferParams& prm, std::vector<SubgroupResult>& out,
                 EngineStats& stats, cudaStream_t stream)
{
    require_device();
    out.assign(in.size(), SubgroupResult());
    if (in.empty()) return;
    const auto w0 = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
    Engine E(prm, stream, stats);
    cudaEvent_t e0, e1;
    RAMBL_CUDA(cudaEventCreate(&e0));
    RAMBL_CUDA(cudaEventCreate(&e1));
    RAMBL_CUDA(cudaEventRecord(e0, stream));
    E.start(in);
    const double ms_start = since(w0);
    {
        const unsigned nt = std::min<unsigned>(host_threads(), 32);
        if (E.subs.size() >= 4 && nt > 1) E.workers.reset(new Workers(std::min<size_t>(nt, E.subs.size())));
    }
    // the device-resident walk takes every subgroup it can; the level-synchronous loop below solves the rest
    {
        const char* ev = getenv("RAMBL_WALK");
        const int mode = prm.level_synchronous ? 0 : (ev ? atoi(ev) : walk_mode());
        if (mode != 0)
        {
            const char* enb = getenv("RAMBL_WALK_NB");
            const char* eti = getenv("RAMBL_WALK_TILE");
            const char* ecl = getenv("RAMBL_WALK_CLUSTER");
            const int taken = E.run_walk(enb ? atoi(enb) : walk_blocks(), eti ? atoi(eti) : 0, ecl ? atoi(ecl) : walk_cluster());
            if (getenv("RAMBL_TRACE")) fprintf(stderr, "[rambl] device walk solved %d of %zu subgroups, until %.1f ms\n", taken, E.subs.size(), since(w0));
        }
    }
    E.upload_pool_chars();
    double trace[4] = {0, 0, 0, 0};
    auto for_subs = [&](const std::function<void(size_t)>& fn) {
        if (E.workers) E.workers->run(E.subs.size(), fn);
        else for (size_t i = 0; i < E.subs.size(); ++i) fn(i);
    };
    for (;;)
    {
        bool any = false;
        for (const Sub& s : E.subs) any = any || !s.done;
        if (!any) break;
        const auto t0 = std::chrono::steady_clock::now();
        for_subs([&](size_t i) { if (!E.subs[i].done) E.prepare(E.subs[i]); });
        const auto t1 = std::chrono::steady_clock::now();
        E.pack_step();
        const auto t2 = std::chrono::steady_clock::now();
        const double* al = E.run_step();
        const auto t3 = std::chrono::steady_clock::now();
        for_subs([&](size_t i) {
            Sub& s = E.subs[i];
            if (s.done) return;
            E.advance(s, s.mode == MODE_NONE ? nullptr : al + s.ab_off);
        });
        const auto t4 = std::chrono::steady_clock::now();
        trace[0] += std::chrono::duration<double, std::milli>(t1 - t0).count();
        trace[1] += std::chrono::duration<double, std::milli>(t2 - t1).count();
        trace[2] += std::chrono::duration<double, std::milli>(t3 - t2).count();
        trace[3] += std::chrono::duration<double, std::milli>(t4 - t3).count();
    }
    if (getenv("RAMBL_TRACE"))
        fprintf(stderr, "[rambl] host ms: prepare %.1f pack %.1f device-step %.1f advance %.1f\n", trace[0], trace[1], trace[2], trace[3]);
    for (Sub& s : E.subs)
        if (!s.have_result && !s.failed) s.status = RAMBL_ERR_NO_STRAINS;
    std::vector<std::vector<double>> infer_ab(E.subs.size());
    for (size_t i = 0; i < E.subs.size(); ++i)
        for (const Cand& c : E.subs[i].result) infer_ab[i].push_back(c.ab);
    const double ms_before_assign = since(w0);
    if (prm.assign)
    {
        const double* al = E.assign_step();
        for (Sub& s : E.subs)
            if (s.mode == MODE_ASSIGN)
                for (size_t k = 0; k < s.result.size(); ++k) s.result[k].ab = al[s.ab_off + k];
    }
    else E.flush_inherits();
    const double ms_walk = since(w0);
    RAMBL_CUDA(cudaEventRecord(e1, stream));
    RAMBL_CUDA(cudaStreamSynchronize(stream));
    E.collect_kernel_times();
    const double ms_times = since(w0);
    E.join_side();
    {
        const EngineStats& q = E.side_stats;
        stats.launches += q.launches; stats.level_steps += q.level_steps; stats.draws += q.draws;
        stats.loglik_updates += q.loglik_updates; stats.gibbs_ms += q.gibbs_ms; stats.gibbs_launches += q.gibbs_launches;
        stats.gibbs_bytes += q.gibbs_bytes; stats.h2d_bytes += q.h2d_bytes; stats.d2h_bytes += q.d2h_bytes;
        stats.gibbs_rounds += q.gibbs_rounds; stats.gibbs_passes += q.gibbs_passes;
    }
    // ---- gather
    for (size_t i = 0; i < E.subs.size(); ++i)
    {
        Sub& s = E.subs[i];
        SubgroupResult& r = out[i];
        if (s.side >= 0) { r = std::move(E.side_out[s.side]); continue; }
        r.status = s.status;
        r.draws = s.draws;
        r.levels = s.levels;
        if (s.status != RAMBL_OK) continue;
        std::vector<double> subs_host((size_t)s.slot_cap * 36);
        RAMBL_CUDA(cudaMemcpyAsync(subs_host.data(), s.sub.p, sizeof(double) * subs_host.size(), cudaMemcpyDeviceToHost, stream));
        stats.d2h_bytes += (long long)sizeof(double) * subs_host.size();
        RAMBL_CUDA(cudaStreamSynchronize(stream));
        for (size_t k = 0; k < s.result.size(); ++k)
        {
            StrainResult sr;
            sr.abundance_infer = infer_ab[i][k];
            sr.abundance = s.result[k].ab;
            sr.path = path_of(s, s.result[k].tail);
            memcpy(sr.sub, &subs_host[(size_t)s.result[k].slot * 36], sizeof(double) * 36);
            if (prm.keep_loglik)
            {
                sr.loglik.resize(s.g->n_reads);
                if (s.g->n_reads)
                    RAMBL_CUDA(cudaMemcpyAsync(sr.loglik.data(), s.ll.p + (size_t)s.result[k].slot * s.R,
                                               sizeof(double) * s.g->n_reads, cudaMemcpyDeviceToHost, stream));
            }
            r.strains.push_back(std::move(sr));
        }
        RAMBL_CUDA(cudaStreamSynchronize(stream));
        r.order.resize(r.strains.size());
        for (size_t k = 0; k < r.order.size(); ++k) r.order[k] = (int)k;
        std::sort(r.order.begin(), r.order.end(), [&](int a, int b) { return r.strains[a].abundance > r.strains[b].abundance; });
    }
    RAMBL_CUDA(cudaStreamSynchronize(stream));
    float ms = 0;
    RAMBL_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    stats.gpu_ms += ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (getenv("RAMBL_TRACE"))
        fprintf(stderr, "[rambl] infer wall ms: start %.1f, walk until %.1f, read_assign until %.1f, kernel times until %.1f, gather until %.1f\n",
                ms_start, ms_before_assign, ms_walk, ms_times, since(w0));
}

}  // namespace rambl
