// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin extern "C" shim that is compiled TOGETHER WITH THE UNMODIFIED REFERENCE
// SOURCES where they lie under /root/reference/StrainCall (see oracle/Makefile;
// outputs go to oracle/_ref/ only).  It lets tests and the bench's CPU-baseline leg
// call the real reference classes on in-memory inputs, bypassing the samtools
// based I/O of StrainCall.cpp:
//
//   ref_msa_align   -> MultipleSequenceAlignmentSP<Index2D,SimpleScoreModel,vector,string,char>::align
//                      (MultipleSequenceAlignment.hpp:87-107, MultipleSequenceAlignmentSP.cpp:10-301),
//                      called exactly as PartialOrderGraph::canonize_insert_at_level does
//                      (PartialOrderGraph.cpp:449-455,507,510-519,546).
//   ref_pog_build   -> PartialOrderGraph::PartialOrderGraph(G,R)   (PartialOrderGraph.cpp:61-265)
//   ref_pog_dump    -> flat dump of every node (id,state,level,out,in,sibling,read_pool)
//   ref_pog_edges   -> PartialOrderGraph::output_edge              (PartialOrderGraph.cpp:318-337)
//   ref_infer       -> infer_strains + read_assign + the abundance sort of main()
//                      (NonparametricClustering.cpp:704-708,776-836; StrainCall.cpp:1021-1027)
//
// All results come back as malloc'ed text (free with ref_free) so that any host
// language can parse them.  Floating point values are printed with %.21Lg (enough
// to round-trip x87 long double) AND as a double (%.17g).
#include "PartialOrderGraph.hpp"
#include "MultipleSequenceAlignment.hpp"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>
#include <chrono>
using namespace std;

namespace {
char* dup_text(const string& s)
{
    char* p = (char*)malloc(s.size() + 1);
    memcpy(p, s.c_str(), s.size() + 1);
    return p;
}
string fmt_ld(long double x)
{
    char b[96];
    snprintf(b, sizeof b, "%.21Lg", x);
    return string(b);
}
string fmt_d(long double x)
{
    char b[64];
    snprintf(b, sizeof b, "%.17g", (double)x);
    return string(b);
}
struct Handle
{
    PartialOrderGraph* pog;
    vector<AlignRead> reads;
    string gene;
};
const char* kAlphabet[6] = {"A", "C", "G", "T", "-", "="};

void dump_strains(ostringstream& os, const char* stage, vector<Strain>& strains, int with_loglik)
{
    os << "STAGE " << stage << " " << strains.size() << "\n";
    for (size_t i = 0; i < strains.size(); ++i)
    {
        Strain& s = strains[i];
        os << "STRAIN " << i << " " << fmt_ld(s.abundance) << " " << fmt_d(s.abundance) << " " << fmt_ld(s.Z) << "\n";
        os << "PATH";
        for (auto p = s.path.begin(); p != s.path.end(); ++p) os << " " << (*p)->id;
        os << "\n";
        os << "SEQ " << s.strain_seq() << "\n";
        os << "PLAIN " << s.plain_seq() << "\n";
        os << "SUB";
        for (int a = 0; a < 6; ++a)
            for (int b = 0; b < 6; ++b)
            {
                auto it = s.sub_count.find(Substitution(kAlphabet[a], kAlphabet[b]));
                os << " " << (it == s.sub_count.end() ? string("nan") : fmt_ld(it->second));
            }
        os << "\n";
        if (with_loglik)
        {
            os << "LOGLIK " << s.read_loglik.size();
            for (auto it = s.read_loglik.begin(); it != s.read_loglik.end(); ++it)
                os << " " << it->first << ":" << fmt_ld(it->second);
            os << "\n";
        }
    }
}
}  // namespace

extern "C" {

void ref_free(char* p) { free(p); }

// Progressive sum-of-pairs alignment of n strings; returns "L\n" followed by n rows of
// L characters (row t = MSA::get(t), the canonised form of input t).
char* ref_msa_align(int n, const char* const* seqs)
{
    MultipleSequenceAlignmentSP<Index2D, SimpleScoreModel, vector, string, char> msa;
    MSA<vector, char> result;
    vector<string> data;
    for (int i = 0; i < n; ++i) data.push_back(string(seqs[i]));
    msa.align(data, result);
    ostringstream os;
    os << result.size() << "\n";
    for (int t = 0; t < n; ++t)
    {
        vector<char> res;
        result.get(t, res);
        os << string(res.begin(), res.end()) << "\n";
    }
    return dup_text(os.str());
}

void* ref_pog_build(const char* gene, int nreads, const int* pos, const char* const* cigar,
                    const char* const* seq, const int* cn)
{
    Handle* h = new Handle;
    h->gene = gene;
    for (int i = 0; i < nreads; ++i)
        h->reads.push_back(AlignRead(pos[i], string(cigar[i]), string(seq[i]), string(""), cn[i]));
    h->pog = new PartialOrderGraph(h->gene, h->reads);
    return h;
}

void ref_pog_free(void* hv)
{
    Handle* h = (Handle*)hv;
    delete h->pog;
    delete h;
}

int ref_pog_num_nodes(void* hv) { return ((Handle*)hv)->pog->N; }

// One line per node, in `nodes` order:
//   NODE id align_state label level | OUT ids | IN ids | SIB ids | POOL rid:label:cn ...
char* ref_pog_dump(void* hv)
{
    PartialOrderGraph* g = ((Handle*)hv)->pog;
    ostringstream os;
    os << "NODES " << g->N << "\n";
    for (auto it = g->nodes.begin(); it != g->nodes.end(); ++it)
    {
        PartialOrderGraphNode* u = *it;
        os << "NODE " << u->id << " " << (int)get<0>(u->state) << " " << get<1>(u->state) << " " << u->level;
        os << " | OUT";
        for (auto o = u->out.begin(); o != u->out.end(); ++o) os << " " << (*o)->id;
        os << " | IN";
        for (auto o = u->in.begin(); o != u->in.end(); ++o) os << " " << (*o)->id;
        os << " | SIB";
        for (auto o = u->sibling.begin(); o != u->sibling.end(); ++o) os << " " << (*o)->id;
        os << " | POOL";
        for (auto r = u->read_pool.begin(); r != u->read_pool.end(); ++r)
            os << " " << get<0>(*r) << ":" << get<1>(*r) << ":" << get<2>(*r);
        os << "\n";
    }
    return dup_text(os.str());
}

char* ref_pog_edges(void* hv)
{
    ostringstream os;
    ((Handle*)hv)->pog->output_edge(os);
    return dup_text(os.str());
}

// read_pairs is given as CSR: for unique read u, mates are pair_val[pair_off[u]..pair_off[u+1]).
// Runs infer_strains(strains,read_pairs,n,e,tau,diff); if do_assign, then
// read_assign(strains,reads,read_pairs,n) and the abundance sort of main().
// elapsed_ms[0] = infer, elapsed_ms[1] = read_assign (wall clock, for the CPU baseline).
char* ref_infer(void* hv, int n_uid, const int* pair_off, const int* pair_val, int n, double e, double tau,
                double diff, int do_assign, int with_loglik, double* elapsed_ms)
{
    Handle* h = (Handle*)hv;
    ReadPairs rp;
    for (int u = 0; u < n_uid; ++u)
        rp[u] = vector<int>(pair_val + pair_off[u], pair_val + pair_off[u + 1]);
    vector<Strain> strains;
    ostringstream os;
    // the CLI holds these as float (StrainCall.cpp:84,89,90) and passes them to DoubleL parameters
    float ef = (float)e, tf = (float)tau, df = (float)diff;
    auto t0 = chrono::steady_clock::now();
    h->pog->infer_strains(strains, rp, n, ef, tf, df);
    auto t1 = chrono::steady_clock::now();
    dump_strains(os, "infer", strains, with_loglik);
    double ms_assign = 0;
    if (do_assign)
    {
        auto t2 = chrono::steady_clock::now();
        h->pog->read_assign(strains, h->reads, rp, n);
        auto t3 = chrono::steady_clock::now();
        ms_assign = chrono::duration<double, milli>(t3 - t2).count();
        dump_strains(os, "assign", strains, 0);
        sort(strains.begin(), strains.end(), [](Strain& a, Strain& b) { return a.abundance > b.abundance; });
        dump_strains(os, "final", strains, 0);
    }
    if (elapsed_ms)
    {
        elapsed_ms[0] = chrono::duration<double, milli>(t1 - t0).count();
        elapsed_ms[1] = ms_assign;
    }
    return dup_text(os.str());
}

}  // extern "C"
