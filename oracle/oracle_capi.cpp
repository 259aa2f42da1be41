// oracle/oracle_capi.cpp -- TEST INFRASTRUCTURE ONLY.
// extern "C" face of the CPU restatement, with the same calls and the same text formats as
// oracle/ref_harness.cpp (prefix orc_ instead of ref_), so one Python parser serves both.
#include "oracle.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>

using namespace oracle;

namespace {
char* dup_text(const std::string& s)
{
    char* p = (char*)malloc(s.size() + 1);
    memcpy(p, s.c_str(), s.size() + 1);
    return p;
}
std::string fmt_ld(LD x) { char b[96]; snprintf(b, sizeof b, "%.21Lg", x); return b; }
std::string fmt_d(LD x) { char b[64]; snprintf(b, sizeof b, "%.17g", (double)x); return b; }
struct Handle
{
    Pog pog;
    std::vector<Read> reads;
};
void dump_strains(std::ostringstream& os, const char* stage, const Pog& g, const std::vector<StrainOut>& v, int with_ll)
{
    os << "STAGE " << stage << " " << v.size() << "\n";
    for (size_t i = 0; i < v.size(); ++i)
    {
        const StrainOut& s = v[i];
        os << "STRAIN " << i << " " << fmt_ld(s.abundance) << " " << fmt_d(s.abundance) << " " << fmt_ld(s.Z) << "\n";
        os << "PATH";
        for (int h : s.path) os << " " << g.store[h].id;
        os << "\nSEQ " << strain_seq(g, s) << "\nPLAIN " << plain_seq(g, s) << "\nSUB";
        for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) os << " " << fmt_ld(s.sub[a][b]);
        os << "\n";
        if (with_ll)
        {
            os << "LOGLIK " << s.loglik.size();
            for (auto& kv : s.loglik) os << " " << kv.first << ":" << fmt_ld(kv.second);
            os << "\n";
        }
    }
}
}  // namespace

extern "C" {

void orc_free(char* p) { free(p); }

char* orc_msa_align(int n, const char* const* seqs)
{
    std::vector<std::string> in;
    for (int i = 0; i < n; ++i) in.push_back(seqs[i]);
    std::vector<std::string> rows = msa_sp_align(in);
    std::ostringstream os;
    os << (rows.empty() ? 0 : rows[0].size()) << "\n";
    for (auto& r : rows) os << r << "\n";
    return dup_text(os.str());
}

void* orc_pog_build(const char* gene, int nreads, const int* pos, const char* const* cigar, const char* const* seq,
                    const int* cn)
{
    Handle* h = new Handle;
    for (int i = 0; i < nreads; ++i) h->reads.push_back({pos[i], cigar[i], seq[i], cn[i]});
    h->pog.build(gene, h->reads);
    return h;
}
void orc_pog_free(void* h) { delete (Handle*)h; }
int orc_pog_num_nodes(void* h) { return (int)((Handle*)h)->pog.order.size(); }
char* orc_pog_dump(void* h) { return dup_text(((Handle*)h)->pog.dump()); }
char* orc_pog_edges(void* h) { return dup_text(((Handle*)h)->pog.edges()); }

char* orc_infer(void* hv, int n_uid, const int* pair_off, const int* pair_val, int n, double e, double tau, double diff,
                int do_assign, int with_loglik, double* elapsed_ms)
{
    Handle* h = (Handle*)hv;
    PairTable pt;
    pt.off.assign(pair_off, pair_off + n_uid + 1);
    pt.val.assign(pair_val, pair_val + pair_off[n_uid]);
    std::vector<StrainOut> strains;
    std::ostringstream os;
    float ef = (float)e, tf = (float)tau, df = (float)diff;  // the CLI holds them as float (StrainCall.cpp:84,89,90)
    auto t0 = std::chrono::steady_clock::now();
    infer_strains(h->pog, pt, n, ef, tf, df, strains);
    auto t1 = std::chrono::steady_clock::now();
    dump_strains(os, "infer", h->pog, strains, with_loglik);
    double ms_assign = 0;
    if (do_assign)
    {
        auto t2 = std::chrono::steady_clock::now();
        read_assign(h->pog, h->reads, pt, n, strains);
        auto t3 = std::chrono::steady_clock::now();
        ms_assign = std::chrono::duration<double, std::milli>(t3 - t2).count();
        dump_strains(os, "assign", h->pog, strains, 0);
        std::vector<int> idx(strains.size());
        for (size_t i = 0; i < idx.size(); ++i) idx[i] = (int)i;
        std::sort(idx.begin(), idx.end(), [&](int a, int b) { return strains[a].abundance > strains[b].abundance; });
        std::vector<StrainOut> sorted;
        for (int i : idx) sorted.push_back(strains[i]);
        dump_strains(os, "final", h->pog, sorted, 0);
    }
    if (elapsed_ms)
    {
        elapsed_ms[0] = std::chrono::duration<double, std::milli>(t1 - t0).count();
        elapsed_ms[1] = ms_assign;
    }
    return dup_text(os.str());
}

}  // extern "C"
