// StrainCall -- drop-in command line for the reference binary of the same name
// (/root/reference/StrainCall/StrainCall.cpp).  Same options (StrainCall.cpp:34-154), same samtools
// based input (faidx / view / mpileup through the shell, temporary files in the working directory),
// same FASTA / -G output, so scripts/rambl.py (rambl.py:165-194) can call it unchanged.  What differs:
// all scan windows are put into ONE batch and solved together on the GPU through the C ABI
// (include/rambl_b200.h); there is no CPU path, the program exits non-zero without a device.
//
// The I/O glue below restates the reference's behaviour, quirks included, because the reads that
// reach the graph must be the same reads: window adjustment (StrainCall.cpp:673-783), read cropping
// (291-414), filters, depth down-sampling with std::mt19937(1234) and the AlignRead de-duplication
// order (480-670).
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <iostream>
#include <map>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

#include "rambl_b200.h"

namespace {

struct Options  // sc_parameter, StrainCall.cpp:58-95
{
    std::string gene_file, mapping_file;
    std::vector<std::string> rois;  // -r may be given several times; --roi-file FILE adds one region per line
    int device = -1;                // --device N: the GPU this process uses (default: the current device)
    int window_size = 500, overlap_size = 100;
    float error_rate = 0.01f;
    int mapping_qual = 3, max_ins = 10, read_len = 80;
    bool print_help = false;
    float tau = 0.02f, diff_rate = 0.01f;
    int max_depth = 800;
    bool plot_graph = false;
    std::string dump_inputs;  // --dump-inputs FILE: write what would be handed to the device, then exit
};

void usage()
{
    std::cerr << "StrainCall marker_gene read_mapping\n"
                 "           [-r gn:p0-p1] [-w window_size]\n"
                 "           [-e error_rate] [-q map_qual]\n\n"
                 "Options\n"
                 "-r,--roi           region of interesting, gn is gene name,\n"
                 "                   p0 is starting position, p1 is ending position (inclusive)\n"
                 "-w,--window        the size of scanning window [500]\n"
                 "-o,--overlap       the size of window-window overlap [100]\n"
                 "-e,--error-rate    sequencing error rate [0.01]\n"
                 "-D,--max-depth     downsample data to the specified depth [800]\n"
                 "-q,--map-qual      only include reads with mapping quality >= INT [3]\n"
                 "-I,--max-ins       only include reads with insertions <= INT [10]\n"
                 "-l,--read-len      only include reads with length >=INT [80]\n"
                 "-t,--tau           only include strains with abundance level >=FLT [0.02]\n"
                 "-d,--diff-rate     only include strains with difference rate >=FLT [0.01]\n"
                 "-G,--plot-graph    print graph\n"
                 "   --roi-file FILE one region per line; with this or several -r all regions are solved\n"
                 "                   as one batch on the GPU (output = the separate runs, concatenated)\n"
                 "   --device INT    CUDA device to use\n"
                 "-h,--help          print this message\n\n";
}

bool is_opt(const std::string& a, const char* s, const char* l)
{
    return a == s || a == std::string("--") + l || a == std::string("-") + l;
}

Options parse(int argc, char** argv)
{
    Options o;
    int positional = 0;
    for (int i = 0; i < argc; ++i)
    {
        const std::string a = argv[i];
        if (!a.empty() && a[0] == '-')
        {
            auto next = [&]() -> std::string { return (i + 1 < argc) ? std::string(argv[++i]) : std::string(); };
            if (is_opt(a, "-h", "help")) o.print_help = true;
            else if (is_opt(a, "-r", "roi")) o.rois.push_back(next());
            else if (a == "--roi-file")
            {
                std::ifstream in(next());
                std::string line;
                while (std::getline(in, line))
                {
                    while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
                    if (!line.empty() && line[0] != '#') o.rois.push_back(line);
                }
            }
            else if (a == "--device") o.device = std::stoi(next());
            else if (is_opt(a, "-w", "window")) o.window_size = std::stoi(next());
            else if (is_opt(a, "-e", "error-rate")) o.error_rate = std::stof(next());
            else if (is_opt(a, "-q", "map-qual")) o.mapping_qual = std::stoi(next());
            else if (is_opt(a, "-o", "overlap")) o.overlap_size = std::stoi(next());
            else if (is_opt(a, "-l", "read-len")) o.read_len = std::stoi(next());
            else if (is_opt(a, "-t", "tau")) o.tau = std::stof(next());
            else if (is_opt(a, "-d", "diff-rate")) o.diff_rate = std::stof(next());
            else if (is_opt(a, "-D", "max-depth")) o.max_depth = std::stoi(next());
            else if (is_opt(a, "-I", "max-ins")) o.max_ins = std::stoi(next());
            else if (is_opt(a, "-G", "plot-graph")) o.plot_graph = true;
            else if (a == "--dump-inputs") o.dump_inputs = next();
        }
        else
        {
            if (positional == 0) o.gene_file = a; else o.mapping_file = a;
            ++positional;
        }
    }
    return o;
}

std::string temp_name(const std::string& stem)
{
    static unsigned counter = 0;
    std::ostringstream os;
    os << stem << "_" << (long)time(0) << "_" << (long)getpid() << "_" << counter++;
    return os.str();
}

void shell(const std::string& cmd) { if (system(cmd.c_str()) == -1) { /* like the reference: ignored */ } }

bool exists(const std::string& p)
{
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}

// load_gene_seq, StrainCall.cpp:157-185
std::string fetch_sequence(const std::string& gene_file, const std::string& roi)
{
    const std::string tmp = temp_name(roi);
    shell("samtools faidx " + gene_file + " " + roi + " 2>/dev/null 1>" + tmp);
    std::string seq, line;
    std::ifstream in(tmp);
    while (std::getline(in, line)) if (!line.empty() && line[0] != '>') seq += line;
    in.close();
    unlink(tmp.c_str());
    return seq;
}

void ensure_index(const std::string& gene_file)
{
    if (!exists(gene_file)) shell("samtools faidx " + gene_file + " 2>/dev/null");  // sic, StrainCall.cpp:227-231
}

// gene_name / gene_length, StrainCall.cpp:222-273
std::string last_gene_name(const std::string& gene_file)
{
    ensure_index(gene_file);
    std::ifstream in(gene_file + ".fai");
    std::string line, name;
    while (std::getline(in, line))
    {
        std::stringstream ss(line);
        std::string f1;
        ss >> f1;
        if (!f1.empty()) name = f1;
    }
    return name;
}

int gene_length(const std::string& gene_file, const std::string& name)
{
    ensure_index(gene_file);
    std::ifstream in(gene_file + ".fai");
    std::string line;
    int len = 0;
    while (std::getline(in, line))
    {
        std::stringstream ss(line);
        std::string f1, f2;
        ss >> f1 >> f2;
        if (f1 == name) len = std::stoi(f2);
    }
    return len;
}

std::string roi_name(const std::string& roi) { return roi.substr(0, roi.find_first_of(':')); }
int roi_start(const std::string& roi)
{
    const size_t x = roi.find_first_of(':');
    std::string num;
    for (size_t i = x + 1; i < roi.size() && roi[i] != '-'; ++i) num.push_back(roi[i]);
    return std::stoi(num);
}
int roi_end(const std::string& roi) { return std::stoi(roi.substr(roi.find_first_of('-') + 1)); }

typedef std::pair<char, int> CigarOp;

// parse_cigar, PartialOrderGraph.cpp:13-59
std::vector<CigarOp> parse_cigar(const std::string& c)
{
    std::vector<CigarOp> r;
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

// window_adjust, StrainCall.cpp:673-783: move the window borders off positions that carry indels
void adjust_window(const Options& o, const std::string& gn, int p0, int p1, int z, int L, int& d0, int& d1)
{
    if (p0 - z < 1) z = p0 - 1;
    int P = p0 - z, Q = std::min(p1 + z, L);
    const std::string tmp = temp_name(gn + ":" + std::to_string(p0) + "-" + std::to_string(p1));
    shell("samtools mpileup -q " + std::to_string(o.mapping_qual) + " -Q0  -A  -r " + gn + ":" + std::to_string(P) + "-" +
          std::to_string(Q) + " " + o.mapping_file + " 2>/dev/null 1>" + tmp);
    std::map<int, std::pair<bool, bool>> info;  // position -> (insertion seen, deletion seen)
    std::ifstream in(tmp);
    std::string line;
    while (std::getline(in, line))
    {
        std::stringstream ss(line);
        std::string f1, f2, f3, f4, f5;
        ss >> f1 >> f2 >> f3 >> f4 >> f5;
        if (f2.empty()) continue;
        const bool ins = f5.find('+') != std::string::npos;
        const bool del = f5.find('-') != std::string::npos || f5.find('*') != std::string::npos;
        info[std::stoi(f2)] = {ins, del};
    }
    in.close();
    unlink(tmp.c_str());
    d0 = d1 = 0;
    if (info.empty()) return;  // the reference dereferences begin() of an empty map here
    auto it0 = info.find(p0);
    if (it0 == info.end()) P = info.begin()->first;
    else
    {
        P = p0;
        while (it0->second.first || it0->second.second)
        {
            if (it0 == info.begin()) break;  // the reference steps before begin() here
            --it0;
            --P;
        }
    }
    auto it1 = info.find(p1);
    if (it1 == info.end()) Q = info.rbegin()->first;
    else
    {
        Q = p1;
        while (it1->second.first || it1->second.second)
        {
            ++it1;
            if (it1 == info.end()) break;
            ++Q;
        }
    }
    d0 = p0 - P;
    d1 = Q - p1;
}

struct Window { std::string gn; int p0, p1; };

// make_scan_window, StrainCall.cpp:798-848
std::vector<Window> scan_windows(const Options& o, const std::string& roi)
{
    std::vector<Window> w;
    std::string gn;
    int l, L, LL;
    if (roi.empty())
    {
        gn = last_gene_name(o.gene_file);
        l = 1;
        L = LL = gene_length(o.gene_file, gn);
    }
    else
    {
        gn = roi_name(roi);
        l = roi_start(roi);
        L = roi_end(roi);
        LL = gene_length(o.gene_file, gn);
    }
    std::set<int> seen;
    int d0 = 0, d1 = 0;
    for (int p0 = l, p1 = l; p1 < L; p0 += o.window_size - o.overlap_size)
    {
        p1 = std::min(p0 + o.window_size - 1, L);
        adjust_window(o, gn, p0, p1, 50, LL, d0, d1);
        if (seen.count(p1 + d1)) continue;
        w.push_back({gn, p0 - d0, p1 + d1});
        seen.insert(p1 + d1);
    }
    return w;
}

// crop_read_within_window, StrainCall.cpp:291-414 (the quality string is never used downstream)
void crop_to_window(int wp0, int wp1, const std::string& seq, const std::vector<CigarOp>& cig, int rp0, int rp1,
                    std::string& out_seq, std::string& out_cigar)
{
    int i = 0, j = 0, ki = 0, kj = 0;
    std::vector<CigarOp> kept;
    size_t a = 0;
    if (cig[a].first == 'S') { i += cig[a].second; ++a; }
    char op = cig[a].first;
    int len = cig[a].second;
    if (rp0 < wp0 && rp0 < wp1)
    {
        while (rp0 < wp0 && rp0 < wp1)
        {
            ki = 0;
            op = cig[a].first;
            len = cig[a].second;
            if (op == 'M') { for (; ki < len; ++ki, ++i, ++rp0) if (rp0 == wp0) break; }
            else if (op == 'D') { for (; ki < len; ++ki, ++rp0) if (rp0 == wp0) break; }
            else if (op == 'I') i += len;
            ++a;
        }
    }
    else ++a;
    if (ki < len) kept.push_back({op, len - ki});
    for (; a < cig.size(); ++a) kept.push_back(cig[a]);

    size_t b = cig.size();  // walks from the back: cig[b-1]
    if (cig[b - 1].first == 'S') { j += cig[b - 1].second; --b; kept.pop_back(); }
    while (rp1 > wp1 && rp1 > wp0)
    {
        kj = 0;
        op = cig[b - 1].first;
        len = cig[b - 1].second;
        if (op == 'M') { for (; kj < len; ++kj, ++j, --rp1) if (rp1 == wp1) break; }
        else if (op == 'D') { for (; kj < len; ++kj, --rp1) if (rp1 == wp1) break; }
        else if (op == 'I') j += len;
        --b;
        if (kj == len || op == 'I') kept.pop_back();
        else kept.back().second -= kj;
    }
    out_seq = seq.substr(i, seq.length() - i - j);
    out_cigar.clear();
    for (const CigarOp& c : kept) out_cigar += std::to_string(c.second) + std::string(1, c.first);
}

struct WindowReads
{
    std::string gene;
    std::vector<int32_t> pos, cn, pair_off, pair_val;
    std::vector<std::string> cigar, seq;
};

// load_mapping_reads, StrainCall.cpp:480-670
void load_reads(const Options& o, const std::string& roi, WindowReads& out)
{
    const std::string tmp = temp_name(roi);
    shell("samtools view " + o.mapping_file + " -q " + std::to_string(o.mapping_qual) + " -F 1804 " + roi +
          " 2>/dev/null 1>" + tmp);
    const int p0 = roi_start(roi), p1 = roi_end(roi);
    std::vector<std::vector<std::string>> recs;
    {
        std::ifstream in(tmp);
        std::string line;
        while (std::getline(in, line))
        {
            std::stringstream ss(line);
            std::vector<std::string> f(11);
            for (int k = 0; k < 11; ++k) ss >> f[k];
            recs.push_back(f);
        }
    }
    unlink(tmp.c_str());
    int depth = 0;
    for (const auto& f : recs)
    {
        int len = 0;
        for (const CigarOp& c : parse_cigar(f[5])) if (c.first == 'M' || c.first == 'D') len += c.second;
        const int r0 = std::stoi(f[3]), r1 = r0 + len - 1;
        if (p0 <= r0 && p1 > r1) depth += r1 - r0 + 1;
        else if (p0 <= r0 && p1 <= r1) depth += p1 - r0 + 1;
        else if (p0 > r0 && p1 <= r1) depth += p1 - p0 + 1;
        else if (p0 > r0 && p1 > r1) depth += r1 - p0 + 1;
    }
    depth /= p1 - p0 + 1;
    const long double rho = std::min(1.0, o.max_depth / (depth + 0.));
    std::mt19937 gen(1234);
    std::uniform_real_distribution<> dicer(0, 1);

    typedef std::tuple<int, std::string, std::string, std::string, int> Key;  // AlignRead ordering
    std::map<Key, std::vector<std::string>> groups;
    for (const auto& f : recs)
    {
        if ((int)f[9].length() < o.read_len) continue;
        if (f[9].find('N') != std::string::npos || f[9].find('n') != std::string::npos) continue;
        std::string name = f[0];
        const int flag = std::stoi(f[1]);
        if ((flag & 65) == 65) name += "/1";
        else if ((flag & 129) == 129) name += "/2";
        const std::vector<CigarOp> cig = parse_cigar(f[5]);
        const int r0 = std::stoi(f[3]);
        int r1 = r0;
        for (const CigarOp& c : cig) if (c.first == 'M' || c.first == 'D') r1 += c.second;
        r1 -= 1;
        const int rel = std::max(0, r0 - p0);
        std::string seq, cigar;
        crop_to_window(p0, p1, f[9], cig, r0, r1, seq, cigar);
        int maxins = 0;
        for (const CigarOp& c : parse_cigar(cigar)) if (c.first == 'I' && c.second > maxins) maxins = c.second;
        if ((int)seq.length() > o.read_len && maxins < o.max_ins)
        {
            if (dicer(gen) > rho) continue;
            groups[Key(rel, cigar, seq, "", 1)].push_back(name);
        }
    }
    std::map<std::string, int> uid_of;
    int id = 0;
    for (auto it = groups.begin(); it != groups.end(); ++it, ++id)
    {
        out.pos.push_back(std::get<0>(it->first));
        out.cigar.push_back(std::get<1>(it->first));
        out.seq.push_back(std::get<2>(it->first));
        out.cn.push_back((int32_t)it->second.size());
        for (const std::string& n : it->second) uid_of[n] = id;
    }
    std::vector<std::vector<int32_t>> mates(out.pos.size());
    for (auto it = uid_of.begin(); it != uid_of.end(); ++it)
    {
        const std::string& n = it->first;
        std::string other;
        if (n.size() >= 2 && n.compare(n.size() - 2, 2, "/1") == 0) other = n.substr(0, n.size() - 2) + "/2";
        else if (n.size() >= 2 && n.compare(n.size() - 2, 2, "/2") == 0) other = n.substr(0, n.size() - 2) + "/1";
        int m = -1;
        if (!other.empty())
        {
            auto jt = uid_of.find(other);
            if (jt != uid_of.end()) m = jt->second;
        }
        mates[it->second].push_back(m);
    }
    out.pair_off.assign(1, 0);
    for (const auto& v : mates)
    {
        out.pair_val.insert(out.pair_val.end(), v.begin(), v.end());
        out.pair_off.push_back((int32_t)out.pair_val.size());
    }
}

}  // namespace

int main(int argc, char** argv)
{
    const Options o = parse(argc - 1, argv + 1);
    if (o.print_help || argc <= 1)
    {
        usage();
        return 0;
    }
    if (o.dump_inputs.empty() && rambl_device_count() < 1)
    {
        std::cerr << "StrainCall (rambl_b200): no CUDA device; this build has no CPU path" << std::endl;
        return 2;
    }
    if (o.device >= 0 && o.dump_inputs.empty() && rambl_set_device(o.device) != RAMBL_OK)
    {
        std::cerr << "StrainCall: " << rambl_last_error() << std::endl;
        return 2;
    }
    // every region of interest (the reference takes one; scripts/rambl.py starts one process per seed gene,
    // rambl.py:179-190) -- the windows of all regions go into ONE batch, and the output is what the separate runs
    // would print one after the other
    std::vector<Window> windows;
    if (o.rois.empty()) windows = scan_windows(o, "");
    for (const std::string& roi : o.rois)
    {
        const std::vector<Window> w = scan_windows(o, roi);
        windows.insert(windows.end(), w.begin(), w.end());
    }
    if (!o.dump_inputs.empty())
    {   // host-only: the windows and the reads exactly as they would enter the graph construction
        std::ofstream out(o.dump_inputs);
        for (const Window& w : windows)
        {
            const std::string roi = w.gn + ":" + std::to_string(w.p0) + "-" + std::to_string(w.p1);
            WindowReads wr;
            wr.gene = fetch_sequence(o.gene_file, roi);
            load_reads(o, roi, wr);
            out << "WINDOW " << w.gn << " " << w.p0 << " " << w.p1 << " " << wr.pos.size() << "\nGENE " << wr.gene << "\n";
            for (size_t i = 0; i < wr.pos.size(); ++i)
            {
                out << "READ " << wr.pos[i] << " " << wr.cigar[i] << " " << wr.seq[i] << " " << wr.cn[i];
                for (int k = wr.pair_off[i]; k < wr.pair_off[i + 1]; ++k) out << " " << wr.pair_val[k];
                out << "\n";
            }
        }
        return 0;
    }
    rambl_batch* b = rambl_batch_create();
    std::vector<Window> kept;
    for (const Window& w : windows)
    {
        const std::string roi = w.gn + ":" + std::to_string(w.p0) + "-" + std::to_string(w.p1);
        WindowReads wr;
        wr.gene = fetch_sequence(o.gene_file, roi);
        load_reads(o, roi, wr);
        if (wr.pos.empty()) continue;  // StrainCall.cpp:1009-1012
        std::vector<const char*> cg, sq;
        for (size_t i = 0; i < wr.pos.size(); ++i) { cg.push_back(wr.cigar[i].c_str()); sq.push_back(wr.seq[i].c_str()); }
        const int sg = rambl_batch_add_subgroup(b, wr.gene.c_str(), (int32_t)wr.pos.size(), wr.pos.data(), cg.data(), sq.data(),
                                                wr.cn.data(), wr.pair_off.data(), wr.pair_val.data());
        if (sg < 0) { std::cerr << "StrainCall: " << rambl_last_error() << std::endl; return 1; }
        kept.push_back(w);
    }
    if (!kept.empty())
    {
        // graphs only for -G; else graphs and strains in one call (construction of one chunk of windows overlaps the
        // strain search of the previous one)
        const int rc = o.plot_graph ? rambl_batch_build_graphs(b) : rambl_batch_solve(b, 5000, o.error_rate, o.tau, o.diff_rate, 1, 0);
        if (rc != RAMBL_OK) { std::cerr << "StrainCall: " << rambl_last_error() << std::endl; return 1; }
    }
    for (size_t i = 0; i < kept.size(); ++i)
    {
        char* txt = nullptr;
        if (o.plot_graph) txt = rambl_batch_graph_text(b, (int32_t)i, 1);
        else if (rambl_batch_status(b, (int32_t)i) == RAMBL_OK)
            txt = rambl_batch_fasta(b, (int32_t)i, kept[i].gn.c_str(), kept[i].p0, kept[i].p1, o.tau);
        else std::cerr << "StrainCall: window " << kept[i].gn << ":" << kept[i].p0 << "-" << kept[i].p1
                       << " has no surviving strain (status " << rambl_batch_status(b, (int32_t)i) << ")" << std::endl;
        if (txt) { std::cout << txt; rambl_free(txt); }
    }
    rambl_batch_destroy(b);
    return 0;
}
