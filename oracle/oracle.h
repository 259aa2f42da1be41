// oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
// Shared declarations of the CPU restatement of RAMBL's StrainCall hot path.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
#pragma once
#include <map>
#include <string>
#include <vector>

namespace oracle {

typedef long double LD;  // the reference's DoubleL (PartialOrderGraph.hpp:229)

enum { ST_MAT = 0, ST_MIS = 1, ST_INS = 2, ST_DEL = 3 };  // AlignState, PartialOrderGraph.hpp:82

struct PoolEntry  // ReadBase = <read id, letters of the read at this node, copy number>
{
    int rid;
    std::string s;
    int cn;
    bool operator<(const PoolEntry& o) const
    {
        if (rid != o.rid) return rid < o.rid;
        if (s != o.s) return s < o.s;
        return cn < o.cn;
    }
};

struct Node
{
    int id = -1;  // position in Pog::order (kept in step by drop_node like the reference does)
    int st = ST_MAT;
    std::string label;
    int level = -1;
    std::vector<int> out, in, sib;  // handles (indices into Pog::store)
    std::vector<PoolEntry> pool;
};

struct Read
{
    int pos;
    std::string cigar, seq;
    int cn;
};

struct Pog
{
    std::vector<Node> store;  // handle -> node (never shrinks)
    std::vector<int> order;   // the reference's `nodes` vector, as handles
    void build(const std::string& gene, const std::vector<Read>& reads);
    std::string dump() const;   // same text as oracle/ref_harness.cpp:ref_pog_dump
    std::string edges() const;  // same text as PartialOrderGraph::output_edge
    int cover(int hu, int hv) const;

    // pieces, named after what they do
    int new_node(int st, const std::string& label);
    void link(int u, int w);
    void unlink(int u, int v);
    bool linked(int u, int v) const;
    void link_chain(int u, const std::vector<int>& gap);
    void link_chain(int u, int v, const std::vector<int>& gap);
    void drop_node(int w, bool bridging);
    void fuse(int u, int v);
};

struct StrainOut
{
    LD abundance = 0, Z = 0;
    std::vector<int> path;  // handles
    LD sub[6][6];
    std::map<int, LD> loglik;
};

// streaming_clustering / read_assign restatement (oracle_dpm.cpp)
struct PairTable
{
    std::vector<int> off, val;  // CSR: mates of unique read u are val[off[u]..off[u+1])
};
void infer_strains(const Pog& g, const PairTable& pairs, int n, LD e, LD tau, LD diff, std::vector<StrainOut>& out);
void read_assign(const Pog& g, const std::vector<Read>& reads, const PairTable& pairs, int n,
                 std::vector<StrainOut>& strains);
std::string strain_seq(const Pog& g, const StrainOut& s);
std::string plain_seq(const Pog& g, const StrainOut& s);

int sp_score(char x, char y);
std::vector<std::string> msa_sp_align(const std::vector<std::string>& seqs);

}  // namespace oracle
