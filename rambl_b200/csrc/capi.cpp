// extern "C" boundary of librambl_b200.so (declarations and reference citations: include/rambl_b200.h).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <chrono>
#include <sstream>
#include <thread>
#include <atomic>
#include <mutex>

#include "dpm.cuh"
#include "engine.hpp"
#include "msa_sp.hpp"
#include "pog.hpp"

using namespace rambl;

namespace {

thread_local std::string g_error;

int fail(int code, const std::string& msg)
{
    g_error = msg;
    return code;
}

template <typename F>
int guarded(F&& f)
{
    try
    {
        f();
        return RAMBL_OK;
    }
    catch (const Error& e) { return fail(e.code, e.what()); }
    catch (const std::bad_alloc&) { return fail(RAMBL_ERR_CAPACITY, "out of host memory"); }
    catch (const std::exception& e) { return fail(RAMBL_ERR_INVALID, e.what()); }
}

char* dup_text(const std::string& s)
{
    char* p = (char*)malloc(s.size() + 1);
    if (p) memcpy(p, s.c_str(), s.size() + 1);
    return p;
}

std::string fmt(double x)
{
    char b[64];
    snprintf(b, sizeof b, "%.17g", x);
    return b;
}

struct Subgroup
{
    std::string gene;
    ReadSet reads;
    SubgroupInput input;
    std::unique_ptr<GraphBuilder> builder;
    FlatGraph graph;
    bool built = false;
    SubgroupResult result;
    bool inferred = false;
};

}  // namespace

struct rambl_batch
{
    std::vector<std::unique_ptr<Subgroup>> subs;
    MsaBatch msa;
    bool threaded = false;
    size_t threaded_upto = 0;
    rambl_stats stats;
    InferParams last;
    rambl_batch() { memset(&stats, 0, sizeof stats); }
};

namespace {

// run fn(i) for i in [lo, hi) on the host cores (subgroups are independent); the first error wins
template <typename F>
void parallel_for(size_t lo, size_t hi, F&& fn)
{
    const size_t n = hi > lo ? hi - lo : 0;
    unsigned nt = (unsigned)std::min<size_t>(std::min<size_t>(host_threads(), 64), n);
    if (nt <= 1)
    {
        for (size_t i = lo; i < hi; ++i) fn(i);
        return;
    }
    std::atomic<size_t> next(lo);
    std::atomic<bool> failed(false);
    std::string what;
    int code = RAMBL_OK;
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; ++t)
        pool.emplace_back([&] {
            for (;;)
            {
                const size_t i = next.fetch_add(1);
                if (i >= hi || failed.load()) return;
                try { fn(i); }
                catch (const Error& e)
                {
                    if (!failed.exchange(true)) { what = e.what(); code = e.code; }
                }
                catch (const std::exception& e)
                {
                    if (!failed.exchange(true)) { what = e.what(); code = RAMBL_ERR_INVALID; }
                }
            }
        });
    for (auto& t : pool) t.join();
    if (failed.load()) throw Error(code, what);
}

// phase A for subgroups [lo, hi): every subgroup lists its alignment problems privately, the lists are then appended to
// `into` in subgroup order.  A subgroup without alignment problems (the common case when the strains differ by
// substitutions) needs nothing from the device: it is finished at once, on the same worker -- its builder's memory is
// still in cache and goes back to the allocator for the next subgroup, instead of 5 MB per subgroup lying cold until a
// second pass.
void thread_range(rambl_batch* b, size_t lo, size_t hi, MsaBatch& into)
{
    std::vector<MsaBatch> local(hi - lo);
    static const MsaResult no_rows;
    parallel_for(lo, hi, [&](size_t i) {
        Subgroup& s = *b->subs[i];
        if (s.built) return;
        s.builder.reset(new GraphBuilder);
        s.builder->thread(s.gene, s.reads, local[i - lo]);
        if (s.builder->n_problems() == 0)
        {
            s.builder->finish(no_rows, s.graph);
            s.builder.reset();
            s.input.graph = &s.graph;
            s.built = true;
        }
    });
    for (size_t i = lo; i < hi; ++i)
    {
        const MsaBatch& m = local[i - lo];
        if (!b->subs[i]->builder) continue;
        b->subs[i]->builder->rebase_problems(into.problems());
        for (int p = 0; p < m.problems(); ++p)
        {
            for (int q = m.prob_seq_off[p]; q < m.prob_seq_off[p + 1]; ++q)
                into.add_sequence(m.chars.data() + m.seq_off[q], m.seq_off[q + 1] - m.seq_off[q]);
            into.end_problem();
        }
    }
}

// phase C for the subgroups of [lo, hi) that wait for aligned rows
void finish_range(rambl_batch* b, size_t lo, size_t hi, const MsaResult& rows)
{
    parallel_for(lo, hi, [&](size_t i) {
        Subgroup& s = *b->subs[i];
        if (s.built || !s.builder) return;
        s.builder->finish(rows, s.graph);
        s.builder.reset();
        s.input.graph = &s.graph;
        s.built = true;
    });
}

// phase A for the subgroups added since the last call
void thread_pending(rambl_batch* b)
{
    thread_range(b, b->threaded_upto, b->subs.size(), b->msa);
    b->threaded_upto = b->subs.size();
}

void finish_pending(rambl_batch* b, const MsaResult& rows)
{
    finish_range(b, 0, b->subs.size(), rows);
    b->msa = MsaBatch();
}

const Subgroup* sub_at(const rambl_batch* b, int sg)
{
    if (!b || sg < 0 || sg >= (int)b->subs.size()) throw Error(RAMBL_ERR_INVALID, "subgroup index out of range");
    return b->subs[sg].get();
}

const StrainResult& strain_at(const rambl_batch* b, int sg, int k)
{
    const Subgroup* s = sub_at(b, sg);
    if (!s->inferred) throw Error(RAMBL_ERR_STATE, "rambl_batch_infer has not run");
    if (k < 0 || k >= (int)s->result.strains.size()) throw Error(RAMBL_ERR_INVALID, "strain index out of range");
    return s->result.strains[k];
}

void strains_block(std::ostringstream& os, const char* stage, const Subgroup& s, const std::vector<int>& order,
                   bool infer_abundance, bool with_ll)
{
    os << "STAGE " << stage << " " << order.size() << "\n";
    for (size_t i = 0; i < order.size(); ++i)
    {
        const StrainResult& r = s.result.strains[order[i]];
        const double ab = infer_abundance ? r.abundance_infer : r.abundance;
        double Z = 0;
        for (int q = 0; q < 36; ++q) Z += r.sub[q];
        os << "STRAIN " << i << " " << fmt(ab) << " " << fmt(ab) << " " << fmt(Z) << "\nPATH";
        for (int u : r.path) os << " " << u;
        os << "\nSEQ " << strain_sequence(s.graph, r.path) << "\nPLAIN " << strain_plain_sequence(s.graph, r.path) << "\nSUB";
        for (int q = 0; q < 36; ++q) os << " " << fmt(r.sub[q]);
        os << "\n";
        if (with_ll && !r.loglik.empty())
        {
            os << "LOGLIK " << r.loglik.size();
            for (size_t q = 0; q < r.loglik.size(); ++q) os << " " << q << ":" << fmt(r.loglik[q]);
            os << "\n";
        }
    }
}

}  // namespace

extern "C" {

const char* rambl_last_error(void) { return g_error.c_str(); }

int rambl_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int rambl_set_device(int32_t device)
{
    return guarded([&] {
        int n = 0;
        RAMBL_CUDA(cudaGetDeviceCount(&n));
        if (device < 0 || device >= n) throw Error(RAMBL_ERR_INVALID, "no such CUDA device");
        RAMBL_CUDA(cudaSetDevice(device));
    });
}

void rambl_free(void* p) { free(p); }

void rambl_release_cached_memory(void) { release_cached_memory(); }
int64_t rambl_cached_host_bytes(void) { return (int64_t)cached_host_bytes(); }

int rambl_set_gibbs_blocks(int32_t blocks)
{
    const int a = blocks < 0 ? -blocks : blocks;
    if (a != 0 && a != 1 && a != 2 && a != 4 && !(blocks == 8)) return RAMBL_ERR_INVALID;
    set_gibbs_blocks(blocks);
    return RAMBL_OK;
}

int rambl_set_walk_mode(int32_t mode)
{
    if (mode != 0 && mode != 1) return RAMBL_ERR_INVALID;
    set_walk_mode(mode);
    return RAMBL_OK;
}

int rambl_set_walk_blocks(int32_t blocks)
{
    if (blocks != 0 && blocks != 1 && blocks != 2 && blocks != 4 && blocks != 8) return RAMBL_ERR_INVALID;
    set_walk_blocks(blocks);
    return RAMBL_OK;
}

int rambl_set_walk_cluster(int32_t ctas)
{
    if (ctas != 0 && ctas != 1 && ctas != 2 && ctas != 4 && ctas != 8) return RAMBL_ERR_INVALID;
    set_walk_cluster(ctas);
    return RAMBL_OK;
}

int rambl_set_host_threads(int32_t n)
{
    if (n < 0) return RAMBL_ERR_INVALID;
    set_host_threads(n);
    return RAMBL_OK;
}

int64_t rambl_msa_rows_capacity(int32_t P, const int32_t* prob_seq_off, const int32_t* seq_off)
{
    int64_t total = 0;
    for (int p = 0; p < P; ++p)
    {
        int64_t letters = 0;
        for (int s = prob_seq_off[p]; s < prob_seq_off[p + 1]; ++s) letters += seq_off[s + 1] - seq_off[s];
        const int64_t cap = std::max<int64_t>(letters, 1);  // a profile is never wider than its letters
        total += cap * (prob_seq_off[p + 1] - prob_seq_off[p]);
    }
    return total;
}

int rambl_msa_sp_align_batch(int32_t P, const int32_t* prob_seq_off, const int32_t* seq_off, const char* letters,
                             int32_t* width, int64_t* row_off, int32_t* row_stride, char* rows, uint64_t* dp_cells,
                             float* kernel_ms)
{
    return guarded([&] {
        if (P < 0 || (P > 0 && (!prob_seq_off || !seq_off || !width || !row_off || !row_stride || !rows)))
            throw Error(RAMBL_ERR_INVALID, "null argument");
        MsaBatch in;
        in.prob_seq_off.assign(prob_seq_off, prob_seq_off + P + 1);
        const int nseq = P ? prob_seq_off[P] : 0;
        in.seq_off.assign(seq_off, seq_off + nseq + 1);
        if (nseq && seq_off[nseq] > 0) in.chars.assign(letters, letters + seq_off[nseq]);
        MsaResult out;
        msa_sp_align_batch(in, out);
        for (int p = 0; p < P; ++p) { width[p] = out.width[p]; row_off[p] = out.row_off[p]; row_stride[p] = out.cap[p]; }
        if (!out.rows.empty()) memcpy(rows, out.rows.data(), out.rows.size());
        if (dp_cells) *dp_cells = out.dp_cells;
        if (kernel_ms) *kernel_ms = out.kernel_ms;
    });
}

rambl_batch* rambl_batch_create(void)
{
    try { return new rambl_batch; }
    catch (...) { return nullptr; }
}

void rambl_batch_destroy(rambl_batch* b) { delete b; }

int rambl_batch_add_subgroup(rambl_batch* b, const char* gene, int32_t n_reads, const int32_t* pos,
                             const char* const* cigar, const char* const* seq, const int32_t* copies,
                             const int32_t* pair_off, const int32_t* pair_val)
{
    int index = -1;
    int rc = guarded([&] {
        if (!b || !gene || n_reads < 0 || (n_reads > 0 && (!pos || !cigar || !seq || !copies)))
            throw Error(RAMBL_ERR_INVALID, "null argument");
        std::unique_ptr<Subgroup> s(new Subgroup);
        s->gene = gene;
        for (int i = 0; i < n_reads; ++i)
        {
            if (copies[i] < 1) throw Error(RAMBL_ERR_INVALID, "copy number must be >= 1");
            if (!cigar[i] || !seq[i]) throw Error(RAMBL_ERR_INVALID, "null read string");
            s->reads.add(pos[i], cigar[i], strlen(cigar[i]), seq[i], strlen(seq[i]), copies[i]);
            s->input.read_cn.push_back(copies[i]);
        }
        if (pair_off && pair_val)
        {
            s->input.pair_off.assign(pair_off, pair_off + n_reads + 1);
            s->input.pair_val.assign(pair_val, pair_val + pair_off[n_reads]);
        }
        else
        {
            s->input.pair_off.assign(1, 0);
            for (int i = 0; i < n_reads; ++i)
            {
                s->input.pair_val.insert(s->input.pair_val.end(), copies[i], -1);
                s->input.pair_off.push_back((int)s->input.pair_val.size());
            }
        }
        b->subs.push_back(std::move(s));
        index = (int)b->subs.size() - 1;
    });
    return rc == RAMBL_OK ? index : -rc;
}

int rambl_batch_add_subgroup_packed(rambl_batch* b, const char* gene, int32_t n_reads, const int32_t* pos,
                                    const int64_t* cigar_off, const char* cigar_chars, const int64_t* seq_off,
                                    const char* seq_chars, const int32_t* copies, const int32_t* pair_off,
                                    const int32_t* pair_val)
{
    int index = -1;
    int rc = guarded([&] {
        if (!b || !gene || n_reads < 0 || (n_reads > 0 && (!pos || !cigar_off || !cigar_chars || !seq_off || !seq_chars || !copies)))
            throw Error(RAMBL_ERR_INVALID, "null argument");
        std::unique_ptr<Subgroup> s(new Subgroup);
        s->gene = gene;
        for (int i = 0; i < n_reads; ++i)
        {
            if (copies[i] < 1) throw Error(RAMBL_ERR_INVALID, "copy number must be >= 1");
            if (cigar_off[i + 1] < cigar_off[i] || seq_off[i + 1] < seq_off[i]) throw Error(RAMBL_ERR_INVALID, "string offsets must not decrease");
        }
        if (n_reads > 0 && (cigar_off[0] < 0 || seq_off[0] < 0)) throw Error(RAMBL_ERR_INVALID, "string offsets must not be negative");
        {   // the arenas are taken over as they are (offsets rebased to the first read's)
            ReadSet& rs = s->reads;
            rs.pos.assign(pos, pos + n_reads);
            rs.cn.assign(copies, copies + n_reads);
            s->input.read_cn.assign(copies, copies + n_reads);
            rs.cigar_off.resize((size_t)n_reads + 1);
            rs.seq_off.resize((size_t)n_reads + 1);
            const int64_t c0 = n_reads ? cigar_off[0] : 0, s0 = n_reads ? seq_off[0] : 0;
            for (int i = 0; i <= n_reads && n_reads; ++i) { rs.cigar_off[i] = cigar_off[i] - c0; rs.seq_off[i] = seq_off[i] - s0; }
            if (n_reads)
            {
                rs.cigar_chars.assign(cigar_chars + c0, cigar_chars + cigar_off[n_reads]);
                rs.seq_chars.assign(seq_chars + s0, seq_chars + seq_off[n_reads]);
            }
        }
        if (pair_off && pair_val)
        {
            s->input.pair_off.assign(pair_off, pair_off + n_reads + 1);
            s->input.pair_val.assign(pair_val, pair_val + pair_off[n_reads]);
        }
        else
        {
            s->input.pair_off.assign(1, 0);
            for (int i = 0; i < n_reads; ++i)
            {
                s->input.pair_val.insert(s->input.pair_val.end(), copies[i], -1);
                s->input.pair_off.push_back((int)s->input.pair_val.size());
            }
        }
        b->subs.push_back(std::move(s));
        index = (int)b->subs.size() - 1;
    });
    return rc == RAMBL_OK ? index : -rc;
}

int rambl_batch_add_graph(rambl_batch* b, int32_t n_nodes, int32_t n_reads, const uint8_t* state,
                          const int32_t* label_off, const char* label_chars, const int32_t* out_off,
                          const int32_t* out_to, const int32_t* pool_off, const int32_t* pool_rid,
                          const int32_t* pool_copies, const int32_t* pool_str_off, const char* pool_chars,
                          const int32_t* read_copies, const int32_t* pair_off, const int32_t* pair_val)
{
    int index = -1;
    int rc = guarded([&] {
        if (!b || n_nodes < 2 || n_reads < 0 || !state || !label_off || !label_chars || !out_off || !pool_off ||
            (n_reads > 0 && !read_copies))
            throw Error(RAMBL_ERR_INVALID, "null or empty graph argument");
        std::unique_ptr<Subgroup> s(new Subgroup);
        FlatGraph& g = s->graph;
        g.n_nodes = n_nodes;
        g.n_reads = n_reads;
        g.st.assign(state, state + n_nodes);
        g.level.assign(n_nodes, -1);
        g.label_off.assign(label_off, label_off + n_nodes + 1);
        g.label_chars.assign(label_chars, label_chars + label_off[n_nodes]);
        g.out_off.assign(out_off, out_off + n_nodes + 1);
        if (out_off[n_nodes]) g.out_to.assign(out_to, out_to + out_off[n_nodes]);
        g.in_off.assign(n_nodes + 1, 0);
        g.pool_off.assign(pool_off, pool_off + n_nodes + 1);
        const int ne = pool_off[n_nodes];
        if (ne)
        {
            if (!pool_rid || !pool_copies || !pool_str_off || !pool_chars) throw Error(RAMBL_ERR_INVALID, "null pool argument");
            g.pool_rid.assign(pool_rid, pool_rid + ne);
            g.pool_cn.assign(pool_copies, pool_copies + ne);
            g.pool_str_off.assign(pool_str_off, pool_str_off + ne + 1);
            g.pool_chars.assign(pool_chars, pool_chars + pool_str_off[ne]);
        }
        else g.pool_str_off.assign(1, 0);
        auto monotone = [&](const auto& off, const char* what) {
            if (off[0] != 0) throw Error(RAMBL_ERR_INVALID, std::string(what) + " must start at 0");
            for (size_t i = 1; i < off.size(); ++i)
                if (off[i] < off[i - 1]) throw Error(RAMBL_ERR_INVALID, std::string(what) + " must not decrease");
        };
        monotone(g.label_off, "label_off");
        monotone(g.out_off, "out_off");
        monotone(g.pool_off, "pool_off");
        monotone(g.pool_str_off, "pool_str_off");
        if (out_off[n_nodes] > 0 && !out_to) throw Error(RAMBL_ERR_INVALID, "out_to is null");
        for (int v : g.out_to) if (v < 0 || v >= n_nodes) throw Error(RAMBL_ERR_INVALID, "edge target out of range");
        for (int r = 0; r < n_reads; ++r)
            if (read_copies[r] < 1) throw Error(RAMBL_ERR_INVALID, "copy number must be >= 1");
        for (int e2 = 0; e2 < ne; ++e2)
        {
            const int r = g.pool_rid[e2];
            if (r < 0 || r >= n_reads) throw Error(RAMBL_ERR_INVALID, "read id out of range");
            // a pool entry draws once per copy and looks its mate up by copy index (NonparametricClustering.cpp:169-189)
            if (g.pool_cn[e2] < 1 || g.pool_cn[e2] > read_copies[r])
                throw Error(RAMBL_ERR_INVALID, "pool_copies must lie in [1, read_copies[read]]");
        }
        for (int u = 0; u < n_nodes; ++u)
            if (g.label_off[u + 1] - g.label_off[u] == 1 && g.label_chars[g.label_off[u]] == '$') g.end_node = u;
        if (g.label(0) != "^" || g.end_node < 0) throw Error(RAMBL_ERR_INVALID, "node 0 must be '^' and one node must be '$'");
        fill_edge_cover(g);
        s->input.read_cn.assign(read_copies, read_copies + n_reads);
        if (pair_off && pair_val)
        {
            s->input.pair_off.assign(pair_off, pair_off + n_reads + 1);
            s->input.pair_val.assign(pair_val, pair_val + pair_off[n_reads]);
        }
        else
        {
            s->input.pair_off.assign(1, 0);
            for (int i = 0; i < n_reads; ++i)
            {
                s->input.pair_val.insert(s->input.pair_val.end(), read_copies[i], -1);
                s->input.pair_off.push_back((int)s->input.pair_val.size());
            }
        }
        s->input.graph = &s->graph;
        s->built = true;
        // keep the bookkeeping of the read-threading phases in step: this subgroup needs neither
        if (b->threaded_upto == b->subs.size()) b->threaded_upto += 1;
        else throw Error(RAMBL_ERR_STATE, "add graphs before adding read subgroups, or build the pending subgroups first");
        b->subs.push_back(std::move(s));
        index = (int)b->subs.size() - 1;
    });
    return rc == RAMBL_OK ? index : -rc;
}

int rambl_batch_thread_reads(rambl_batch* b)
{
    return guarded([&] {
        if (!b) throw Error(RAMBL_ERR_INVALID, "null batch");
        thread_pending(b);
    });
}

char* rambl_batch_msa_problems_text(rambl_batch* b)
{
    if (!b) return nullptr;
    std::ostringstream os;
    const MsaBatch& m = b->msa;
    for (int p = 0; p < m.problems(); ++p)
    {
        os << "P " << p << " " << (m.prob_seq_off[p + 1] - m.prob_seq_off[p]) << "\n";
        for (int s = m.prob_seq_off[p]; s < m.prob_seq_off[p + 1]; ++s)
            os << std::string(m.chars.data() + m.seq_off[s], m.chars.data() + m.seq_off[s + 1]) << "\n";
    }
    return dup_text(os.str());
}

int rambl_batch_finish_graphs_with_rows(rambl_batch* b, const char* rows_text)
{
    return guarded([&] {
        if (!b || !rows_text) throw Error(RAMBL_ERR_INVALID, "null argument");
        thread_pending(b);
        MsaResult rows;
        const int P = b->msa.problems();
        rows.width.assign(P, 0); rows.cap.assign(P, 0); rows.row_off.assign(P + 1, 0);
        std::istringstream is(rows_text);
        std::string line;
        int seen = 0;
        while (std::getline(is, line))
        {
            if (line.empty()) continue;
            int p = -1, n = 0;
            if (sscanf(line.c_str(), "P %d %d", &p, &n) != 2 || p != seen || p >= P)
                throw Error(RAMBL_ERR_INVALID, "rows text: expected 'P <index> <n>' in order");
            if (n != b->msa.prob_seq_off[p + 1] - b->msa.prob_seq_off[p])
                throw Error(RAMBL_ERR_INVALID, "rows text: wrong number of rows for a problem");
            rows.row_off[p] = (long long)rows.rows.size();
            for (int t = 0; t < n; ++t)
            {
                if (!std::getline(is, line)) throw Error(RAMBL_ERR_INVALID, "rows text: truncated");
                if (t == 0) { rows.width[p] = (int)line.size(); rows.cap[p] = (int)line.size(); }
                if ((int)line.size() != rows.width[p]) throw Error(RAMBL_ERR_INVALID, "rows text: ragged rows");
                rows.rows.insert(rows.rows.end(), line.begin(), line.end());
            }
            ++seen;
        }
        if (seen != P) throw Error(RAMBL_ERR_INVALID, "rows text: missing problems");
        rows.row_off[P] = (long long)rows.rows.size();
        finish_pending(b, rows);
    });
}

int rambl_batch_build_graphs(rambl_batch* b)
{
    return guarded([&] {
        if (!b) throw Error(RAMBL_ERR_INVALID, "null batch");
        require_device();
        thread_pending(b);
        MsaResult rows;
        msa_sp_align_batch(b->msa, rows);
        b->stats.gpu_launches += rows.launches;
        b->stats.msa_dp_cells += (int64_t)rows.dp_cells;
        b->stats.msa_problems += b->msa.problems();
        b->stats.msa_kernel_ms += rows.kernel_ms;
        b->stats.h2d_bytes += (int64_t)(b->msa.chars.size() + 4 * (b->msa.seq_off.size() + b->msa.prob_seq_off.size()));
        b->stats.d2h_bytes += (int64_t)rows.rows.size();
        finish_pending(b, rows);
    });
}

int rambl_batch_infer(rambl_batch* b, int32_t n, float e, float tau, float diff, int32_t do_assign, int32_t keep_loglik)
{
    return guarded([&] {
        if (!b) throw Error(RAMBL_ERR_INVALID, "null batch");
        std::vector<SubgroupInput> in;
        for (auto& sp : b->subs)
        {
            if (!sp->built) throw Error(RAMBL_ERR_STATE, "graphs are not built");
            in.push_back(sp->input);
        }
        InferParams prm;
        prm.n = n; prm.e = e; prm.tau = tau; prm.diff = diff; prm.assign = do_assign != 0; prm.keep_loglik = keep_loglik != 0;
        std::vector<SubgroupResult> out;
        EngineStats es;
        const auto w0 = std::chrono::steady_clock::now();
        infer_batch(in, prm, out, es);
        if (getenv("RAMBL_TRACE"))
            fprintf(stderr, "[rambl] rambl_batch_infer: infer_batch returned after %.1f ms\n",
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count());
        for (size_t i = 0; i < out.size(); ++i) { b->subs[i]->result = std::move(out[i]); b->subs[i]->inferred = true; }
        b->stats.gpu_launches += es.launches;
        b->stats.level_steps += es.level_steps;
        b->stats.draws += es.draws;
        b->stats.loglik_updates += es.loglik_updates;
        b->stats.infer_gpu_ms += es.gpu_ms;
        b->stats.h2d_bytes += es.h2d_bytes;
        b->stats.d2h_bytes += es.d2h_bytes;
        b->stats.gibbs_kernel_ms += es.gibbs_ms;
        b->stats.gibbs_launches += es.gibbs_launches;
        b->stats.gibbs_alg_bytes += es.gibbs_bytes;
        b->stats.gibbs_rounds += es.gibbs_rounds;
        b->stats.gibbs_passes += es.gibbs_passes;
        b->stats.dpm_kernel_ms += es.walk_ms;
        b->stats.dpm_launches += es.walk_launches;
        b->stats.dpm_alg_bytes += es.walk_bytes;
        b->stats.offtable_levels += es.offtable_levels;
        b->last = prm;
    });
}

// Chunk bounds of the overlapped solve for a batch of N subgroups on a device with `sms` SMs (bound[c] .. bound[c+1] is
// chunk c).  A batch of up to one wave of the walk kernel is one chunk: it gains nothing from chunks (its chains are
// latency-bound and get clusters instead).  Otherwise the first chunk is one wave minus a few SMs -- a walk CTA takes the
// whole shared memory of its SM for half a second or more, and the small launches of the NEXT chunk's set-up (its
// insertion alignment, the model initialisation) need somewhere to run before the first walk CTAs retire -- and the rest
// follows in chunks of at most eight waves.  $RAMBL_SOLVE_CHUNKS = k: k equal chunks instead; $RAMBL_SOLVE_FIRST = n: size
// of the first chunk (both for measurements and tests; results do not depend on the layout).
static std::vector<size_t> solve_chunk_bounds(size_t N, int sms)
{
    std::vector<size_t> bound(1, 0);
    const size_t wave = (size_t)std::max(sms, 1);
    const char* ef = getenv("RAMBL_SOLVE_FIRST");
    const char* ec = getenv("RAMBL_SOLVE_CHUNKS");
    size_t first = (size_t)std::max(sms - 8, 1);
    if (ef) first = (size_t)std::max(1, atoi(ef));
    if (ec && atoi(ec) >= 1)
    {
        const size_t k = std::min<size_t>((size_t)atoi(ec), std::max<size_t>(N, 1));
        for (size_t c = 1; c <= k; ++c) bound.push_back(N * c / k);
    }
    else if (N <= first || (!ef && N <= wave)) bound.push_back(N);
    else
    {
        bound.push_back(first);
        const size_t rest = N - first, cap = 8 * wave;
        const size_t k = (rest + cap - 1) / cap;
        for (size_t c = 1; c <= k; ++c) bound.push_back(first + rest * c / k);
    }
    return bound;
}

int32_t rambl_solve_layout(int32_t n_subgroups, int32_t sms, int32_t* bounds, int32_t cap)
{
    if (n_subgroups < 0 || sms < 1 || (cap > 0 && !bounds)) return -fail(RAMBL_ERR_INVALID, "bad layout query");
    const std::vector<size_t> b = solve_chunk_bounds((size_t)n_subgroups, sms);
    for (size_t i = 0; i < b.size() && (int32_t)i < cap; ++i) bounds[i] = (int32_t)b[i];
    return (int32_t)b.size();
}

// rambl_batch_build_graphs + rambl_batch_infer as ONE call that overlaps them.  The subgroups are dealt into chunks --
// the first is one wave of the walk kernel (one subgroup per SM), so that the device starts as early as possible; the rest
// follows in one chunk (several for very large batches) -- and two driver threads, each with its own CUDA stream, take the
// chunks in order.  The host-heavy part of a chunk (graph construction and the level tables, on all worker threads) is
// done by one driver at a time; it ends when the chunk's walk kernel is in its stream, and the other driver starts on the
// next chunk while that kernel runs.  Kernels of consecutive chunks share the SMs: the later one fills in as the CTAs of
// the earlier one retire.  Results are what the two separate calls give (subgroups never interact).
int rambl_batch_solve(rambl_batch* b, int32_t n, float e, float tau, float diff, int32_t do_assign, int32_t keep_loglik)
{
    return guarded([&] {
        if (!b) throw Error(RAMBL_ERR_INVALID, "null batch");
        require_device();
        const size_t N = b->subs.size();
        bool fresh = b->threaded_upto == 0 && b->msa.problems() == 0;
        for (auto& sp : b->subs) fresh = fresh && !sp->built;
        int device = 0, sms = 148;
        RAMBL_CUDA(cudaGetDevice(&device));
        RAMBL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        const std::vector<size_t> bound = solve_chunk_bounds(N, sms);
        size_t n_drivers = 2;
        if (const char* ev = getenv("RAMBL_SOLVE_DRIVERS")) n_drivers = (size_t)std::max(1, atoi(ev));
        const size_t n_chunks = bound.size() - 1;
        n_drivers = std::min(n_drivers, std::max<size_t>(n_chunks, 1));
        if (!fresh || n_chunks <= 1)
        {   // nothing to overlap (or a batch that is partly built already): the two calls, one after the other
            int rc = rambl_batch_build_graphs(b);
            if (rc != RAMBL_OK) throw Error(rc, g_error);
            rc = rambl_batch_infer(b, n, e, tau, diff, do_assign, keep_loglik);
            if (rc != RAMBL_OK) throw Error(rc, g_error);
            return;
        }
        InferParams prm;
        prm.n = n; prm.e = e; prm.tau = tau; prm.diff = diff; prm.assign = do_assign != 0; prm.keep_loglik = keep_loglik != 0;
        prm.max_cluster = 1;  // the kernels of consecutive chunks share the SMs: one CTA per subgroup
        const auto w0 = std::chrono::steady_clock::now();
        size_t next = 0;      // guarded by host_mu
        std::mutex mu, host_mu;
        std::string what;
        int code = RAMBL_OK;
        auto drive = [&] {
            cudaStream_t st = nullptr;
            try
            {
                RAMBL_CUDA(cudaSetDevice(device));
                RAMBL_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
                for (;;)
                {
                    // ---- the host-heavy phase of one chunk at a time, chunks in order
                    std::unique_lock<std::mutex> host_phase(host_mu);
                    const size_t c = next;
                    {
                        std::lock_guard<std::mutex> lk(mu);
                        if (c >= n_chunks || code != RAMBL_OK) break;
                    }
                    next += 1;
                    const size_t lo = bound[c], hi = bound[c + 1];
                    // graphs of this chunk: splice (and finish what needs no alignment), align the insertion levels of the
                    // others on the device, finish those
                    MsaBatch msa;
                    MsaResult rows;
                    thread_range(b, lo, hi, msa);
                    if (msa.problems() > 0)
                    {
                        msa_sp_align_batch(msa, rows, st);
                        finish_range(b, lo, hi, rows);
                    }
                    // ---- strain search of this chunk; the host phase is over once its walk kernel is launched
                    std::vector<SubgroupInput> in;
                    for (size_t i = lo; i < hi; ++i) in.push_back(b->subs[i]->input);
                    std::vector<SubgroupResult> out;
                    EngineStats es;
                    InferParams mine = prm;
                    mine.on_device_phase = [&host_phase] { if (host_phase.owns_lock()) host_phase.unlock(); };
                    infer_batch(in, mine, out, es, st);
                    if (host_phase.owns_lock()) host_phase.unlock();
                    std::lock_guard<std::mutex> lk(mu);
                    for (size_t i = lo; i < hi; ++i) { b->subs[i]->result = std::move(out[i - lo]); b->subs[i]->inferred = true; }
                    b->stats.gpu_launches += rows.launches + es.launches;
                    b->stats.msa_dp_cells += (int64_t)rows.dp_cells;
                    b->stats.msa_problems += msa.problems();
                    b->stats.msa_kernel_ms += rows.kernel_ms;
                    b->stats.h2d_bytes += (int64_t)(msa.chars.size() + 4 * (msa.seq_off.size() + msa.prob_seq_off.size())) + es.h2d_bytes;
                    b->stats.d2h_bytes += (int64_t)rows.rows.size() + es.d2h_bytes;
                    b->stats.level_steps += es.level_steps;
                    b->stats.draws += es.draws;
                    b->stats.loglik_updates += es.loglik_updates;
                    b->stats.infer_gpu_ms += es.gpu_ms;
                    b->stats.gibbs_kernel_ms += es.gibbs_ms;
                    b->stats.gibbs_launches += es.gibbs_launches;
                    b->stats.gibbs_alg_bytes += es.gibbs_bytes;
                    b->stats.gibbs_rounds += es.gibbs_rounds;
                    b->stats.gibbs_passes += es.gibbs_passes;
                    b->stats.dpm_kernel_ms += es.walk_ms;
                    b->stats.dpm_launches += es.walk_launches;
                    b->stats.dpm_alg_bytes += es.walk_bytes;
                    b->stats.offtable_levels += es.offtable_levels;
                }
            }
            catch (const Error& er)
            {
                std::lock_guard<std::mutex> lk(mu);
                if (code == RAMBL_OK) { code = er.code; what = er.what(); }
            }
            catch (const std::exception& er)
            {
                std::lock_guard<std::mutex> lk(mu);
                if (code == RAMBL_OK) { code = RAMBL_ERR_INVALID; what = er.what(); }
            }
            if (st) cudaStreamDestroy(st);
        };
        std::vector<std::thread> others;
        for (size_t t = 1; t < n_drivers; ++t) others.emplace_back(drive);
        drive();
        for (std::thread& t : others) t.join();
        b->threaded_upto = N;
        b->msa = MsaBatch();
        b->last = prm;
        if (getenv("RAMBL_TRACE"))
            fprintf(stderr, "[rambl] rambl_batch_solve: %zu subgroups in %zu chunks, %.1f ms\n", N, n_chunks,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count());
        if (code != RAMBL_OK) throw Error(code, what);
    });
}

int32_t rambl_batch_num_subgroups(const rambl_batch* b) { return b ? (int32_t)b->subs.size() : 0; }

int32_t rambl_batch_num_nodes(const rambl_batch* b, int32_t sg)
{
    try { const Subgroup* s = sub_at(b, sg); return s->built ? s->graph.n_nodes : -RAMBL_ERR_STATE; }
    catch (const Error& e) { return -fail(e.code, e.what()); }
}

char* rambl_batch_graph_text(const rambl_batch* b, int32_t sg, int32_t what)
{
    try
    {
        const Subgroup* s = sub_at(b, sg);
        if (!s->built) throw Error(RAMBL_ERR_STATE, "graphs are not built");
        return dup_text(what == 0 ? s->graph.dump() : s->graph.edges());
    }
    catch (const Error& e) { fail(e.code, e.what()); return nullptr; }
}

int32_t rambl_batch_status(const rambl_batch* b, int32_t sg)
{
    try { const Subgroup* s = sub_at(b, sg); return s->inferred ? s->result.status : RAMBL_ERR_STATE; }
    catch (const Error& e) { return fail(e.code, e.what()); }
}

int32_t rambl_batch_num_strains(const rambl_batch* b, int32_t sg)
{
    try { const Subgroup* s = sub_at(b, sg); return s->inferred ? (int32_t)s->result.strains.size() : -RAMBL_ERR_STATE; }
    catch (const Error& e) { return -fail(e.code, e.what()); }
}

int rambl_batch_strain(const rambl_batch* b, int32_t sg, int32_t k, double* abundance_infer, double* abundance,
                       int32_t* path_len)
{
    return guarded([&] {
        const StrainResult& r = strain_at(b, sg, k);
        if (abundance_infer) *abundance_infer = r.abundance_infer;
        if (abundance) *abundance = r.abundance;
        if (path_len) *path_len = (int32_t)r.path.size();
    });
}

int rambl_batch_strain_path(const rambl_batch* b, int32_t sg, int32_t k, int32_t* path)
{
    return guarded([&] {
        const StrainResult& r = strain_at(b, sg, k);
        for (size_t i = 0; i < r.path.size(); ++i) path[i] = r.path[i];
    });
}

char* rambl_batch_strain_sequence(const rambl_batch* b, int32_t sg, int32_t k, int32_t plain)
{
    try
    {
        const StrainResult& r = strain_at(b, sg, k);
        const Subgroup* s = sub_at(b, sg);
        return dup_text(plain ? strain_plain_sequence(s->graph, r.path) : strain_sequence(s->graph, r.path));
    }
    catch (const Error& e) { fail(e.code, e.what()); return nullptr; }
}

int rambl_batch_strain_sub(const rambl_batch* b, int32_t sg, int32_t k, double* sub36)
{
    return guarded([&] { memcpy(sub36, strain_at(b, sg, k).sub, sizeof(double) * 36); });
}

int rambl_batch_strain_loglik(const rambl_batch* b, int32_t sg, int32_t k, double* loglik, int32_t n_reads)
{
    return guarded([&] {
        const StrainResult& r = strain_at(b, sg, k);
        if ((int)r.loglik.size() != n_reads) throw Error(RAMBL_ERR_STATE, "log-likelihood rows were not kept (keep_loglik) or size mismatch");
        memcpy(loglik, r.loglik.data(), sizeof(double) * n_reads);
    });
}

int rambl_batch_order(const rambl_batch* b, int32_t sg, int32_t* order)
{
    return guarded([&] {
        const Subgroup* s = sub_at(b, sg);
        if (!s->inferred) throw Error(RAMBL_ERR_STATE, "rambl_batch_infer has not run");
        for (size_t i = 0; i < s->result.order.size(); ++i) order[i] = s->result.order[i];
    });
}

char* rambl_batch_strains_text(const rambl_batch* b, int32_t sg)
{
    try
    {
        const Subgroup* s = sub_at(b, sg);
        if (!s->inferred) throw Error(RAMBL_ERR_STATE, "rambl_batch_infer has not run");
        std::ostringstream os;
        std::vector<int> natural(s->result.strains.size());
        for (size_t i = 0; i < natural.size(); ++i) natural[i] = (int)i;
        strains_block(os, "infer", *s, natural, true, true);
        if (b->last.assign)
        {
            strains_block(os, "assign", *s, natural, false, false);
            strains_block(os, "final", *s, s->result.order, false, false);
        }
        return dup_text(os.str());
    }
    catch (const Error& e) { fail(e.code, e.what()); return nullptr; }
}

char* rambl_batch_fasta(const rambl_batch* b, int32_t sg, const char* gene_name, int32_t p0, int32_t p1, float tau)
{
    try
    {
        const Subgroup* s = sub_at(b, sg);
        if (!s->inferred) throw Error(RAMBL_ERR_STATE, "rambl_batch_infer has not run");
        std::ostringstream os;
        int si = 0;
        for (int k : s->result.order)
        {
            const StrainResult& r = s->result.strains[k];
            // long double comparison against the float parameter in the reference (StrainCall.cpp:1035)
            if (r.abundance >= (double)tau)
                os << ">contig" << (gene_name ? gene_name : "") << p0 << p1 << si << "\n"
                   << strain_plain_sequence(s->graph, r.path) << "\n";
            ++si;
        }
        return dup_text(os.str());
    }
    catch (const Error& e) { fail(e.code, e.what()); return nullptr; }
}

int rambl_batch_walk_plan(const rambl_batch* b, int32_t sg, int64_t out[8])
{
    return guarded([&] {
        if (!out) throw Error(RAMBL_ERR_INVALID, "null argument");
        const Subgroup* s = sub_at(b, sg);
        if (!s->built) throw Error(RAMBL_ERR_STATE, "graphs are not built");
        const WalkEligibility w = walk_eligibility(s->graph, s->input);
        out[0] = w.eligible; out[1] = w.handoff; out[2] = w.reason; out[3] = w.levels;
        out[4] = w.entries; out[5] = w.max_entries; out[6] = w.max_draws; out[7] = w.offtable_levels;
    });
}

int rambl_batch_stats(const rambl_batch* b, rambl_stats* out)
{
    if (!b || !out) return fail(RAMBL_ERR_INVALID, "null argument");
    *out = b->stats;
    return RAMBL_OK;
}

}  // extern "C"
