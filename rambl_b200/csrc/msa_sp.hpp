// Batched sum-of-pairs insertion alignment (device side in msa_sp.cu).
#pragma once
#include "common.hpp"

namespace rambl {

// No compiled-in limit on columns or letters: problems are dealt into shared-memory size classes and anything
// larger runs with its tables in global memory (msa_sp.cu).

// CSR of problems -> sequences -> letters.  Sequences of a problem are aligned in the given order
// (the caller sorts them the way PartialOrderGraph::canonize_insert_at_level does).
struct MsaBatch
{
    std::vector<int> prob_seq_off{0};
    std::vector<int> seq_off{0};
    std::vector<char> chars;
    void begin_problem() {}
    void add_sequence(const char* s, int len)
    {
        chars.insert(chars.end(), s, s + len);
        seq_off.push_back((int)chars.size());
    }
    void end_problem() { prob_seq_off.push_back((int)seq_off.size() - 1); }
    int problems() const { return (int)prob_seq_off.size() - 1; }
};

struct MsaResult
{
    std::vector<int> width;           // final profile width per problem
    std::vector<int> cap;             // row stride per problem
    std::vector<long long> row_off;   // rows of problem p start at rows[row_off[p]], row t at + t*cap[p]
    std::vector<char> rows;
    unsigned long long dp_cells = 0;  // (profile columns x letters) summed over every alignment step
    float kernel_ms = 0;
    int launches = 0;
    std::string row(int p, int t) const
    {
        const char* b = rows.data() + row_off[p] + (long long)t * cap[p];
        return std::string(b, b + width[p]);
    }
};

void msa_sp_align_batch(const MsaBatch& in, MsaResult& out, cudaStream_t stream = 0);

}  // namespace rambl
