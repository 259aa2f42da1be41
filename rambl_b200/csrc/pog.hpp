// Host side of the partial order graph: construction from aligned reads and the flat,
// device-friendly form the clustering engine consumes.
//
// Mirrors the behaviour of the reference's PartialOrderGraph(G,R) constructor
// (/root/reference/StrainCall/PartialOrderGraph.cpp:61-265) -- node order, ordered edge lists and
// ordered read pools are reproduced exactly, because the strain search downstream depends on them --
// but is organised for batching: construction is split in three phases so that the insertion
// alignments of ALL levels of ALL subgroups go to the GPU in one launch (msa_sp.cu):
//   phase A  thread(): splice every read into the backbone by its CIGAR, then list, per backbone
//            level, the insertion strings that need a sum-of-pairs alignment;
//   phase B  msa_sp_align_batch() on the device (not in this file);
//   phase C  finish(): canonise insertions with the aligned rows, canonise deletions, merge equal
//            siblings forward and backward, collapse linear paths, level the nodes, flatten.
#pragma once
#include <cstdint>
#include <string>
#include <string_view>
#include <vector>

#include "msa_sp.hpp"

namespace rambl {

enum : uint8_t { ST_MAT = 0, ST_MIS = 1, ST_INS = 2, ST_DEL = 3 };  // AlignState, PartialOrderGraph.hpp:82

// The aligned reads of one subgroup (AlignRead, PartialOrderGraph.hpp:218; the quality string is never used): positions
// and copy numbers as arrays, CIGARs and letters as two byte arenas with offsets -- the layout the packed C-ABI call
// hands over, so that adding a subgroup is a handful of block copies instead of two strings per read.
struct ReadSet
{
    std::vector<int> pos, cn;
    std::vector<int64_t> cigar_off{0}, seq_off{0};
    std::vector<char> cigar_chars, seq_chars;
    size_t size() const { return pos.size(); }
    std::string_view cigar(size_t i) const { return std::string_view(cigar_chars.data() + cigar_off[i], (size_t)(cigar_off[i + 1] - cigar_off[i])); }
    std::string_view seq(size_t i) const { return std::string_view(seq_chars.data() + seq_off[i], (size_t)(seq_off[i + 1] - seq_off[i])); }
    void add(int p, const char* cigar_text, size_t cigar_len, const char* letters, size_t n_letters, int copies)
    {
        pos.push_back(p);
        cn.push_back(copies);
        cigar_chars.insert(cigar_chars.end(), cigar_text, cigar_text + cigar_len);
        seq_chars.insert(seq_chars.end(), letters, letters + n_letters);
        cigar_off.push_back((int64_t)cigar_chars.size());
        seq_off.push_back((int64_t)seq_chars.size());
    }
};

// The graph as arrays.  Node ids are the reference's ids (index into its `nodes` vector).
struct FlatGraph
{
    int n_nodes = 0;
    int n_reads = 0;                   // unique reads (rid range)
    std::vector<uint8_t> st;           // AlignState per node
    std::vector<int> level;            // LevelOrderIterator level (only output_edge prints it)
    std::vector<int> label_off;        // n_nodes+1, into label_chars
    std::vector<char> label_chars;
    std::vector<int> out_off;          // n_nodes+1, CSR of ordered successors
    std::vector<int> out_to;
    std::vector<int> out_cover;        // number_of_reads_cover_nodes(u, v) per edge
    std::vector<int> in_off;           // ordered predecessors (dump / parity only)
    std::vector<int> in_from;
    std::vector<int> pool_off;         // n_nodes+1, CSR of ordered read-pool entries
    HostVec<int> pool_rid;             // (the four per-entry arrays come from the host block cache, common.hpp)
    HostVec<int> pool_cn;
    HostVec<int> pool_str_off;         // entries+1, into pool_chars
    HostVec<char> pool_chars;
    int end_node = -1;                 // id of "$"

    std::string label(int u) const { return std::string(label_chars.data() + label_off[u], label_chars.data() + label_off[u + 1]); }
    std::string pool_str(int e) const { return std::string(pool_chars.data() + pool_str_off[e], pool_chars.data() + pool_str_off[e + 1]); }
    std::string dump() const;   // NODE lines in the format of oracle/ref_harness.cpp (without SIB)
    std::string edges() const;  // PartialOrderGraph::output_edge
};

// number_of_reads_cover_nodes (PartialOrderGraph.cpp:1218-1244) for every edge of a flat graph whose
// out_cover is still empty (graphs handed in from outside)
void fill_edge_cover(FlatGraph& g);

class GraphBuilder
{
public:
    GraphBuilder();
    ~GraphBuilder();
    GraphBuilder(const GraphBuilder&) = delete;
    GraphBuilder& operator=(const GraphBuilder&) = delete;

    // phase A; appends this graph's alignment problems to `batch` and remembers where they start
    void thread(const std::string& gene, const ReadSet& reads, MsaBatch& batch);
    // phase C; `rows` holds the solved problems of the batch that thread() appended to
    void finish(const MsaResult& rows, FlatGraph& out);
    int n_problems() const;
    // the problems thread() listed start at index `first` of the batch that finish() will be given
    void rebase_problems(int first);

private:
    struct Impl;
    Impl* m;
};

}  // namespace rambl
