// builder.cu -- utree-build_gg on the GPU: FASTA + label map -> .ubt, byte-identical to the
// reference builder (itree.c:501-635 parser + complevel rule, :268-307 xeTreeU_RF relabelling, :1317-1343
// writer, :1225-1232 .log), which is a serial program: 2^24 per-prefix binary trees, one insert per k-mer.
//
// What the serial order decides, and how it is kept without the trees:
//   * a k-mer met again under a different label is relabelled to the longest common prefix of the node's
//     CURRENT label and the new one that ends before a shared ';' and holds >= 2 of them, else it goes bad
//     for good (itree.c:285-305).  The current label may itself be such a prefix, so the result depends on the
//     order of the occurrences: the (word, occurrence) pairs are generated in input order, sorted by word with a
//     STABLE radix sort, and one thread per distinct word folds its occurrences front to back;
//   * label ids are handed out in order of first registration -- a sequence's own label when its header is read
//     (itree.c:594), a derived prefix the first time some fold creates it (itree.c:299), superseded ones
//     included.  Every string that can ever be a label (the map's labels and their cuts before a ';') gets a
//     canonical number up front, the folds record the earliest time each one is created (atomicMin on the byte
//     offset of the k-mer that caused it), and ids are those strings ordered by that time.
// The output then follows: good words in ascending order with their ids, per-label counts, header patched.
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "utb_internal.h"

#define BCK(call)                                                                                                        \
    do {                                                                                                                 \
        cudaError_t e_ = (call);                                                                                         \
        if (e_ != cudaSuccess) {                                                                                         \
            utb_set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
            rc = UTB_ERR_CUDA;                                                                                           \
            goto done;                                                                                                   \
        }                                                                                                                \
    } while (0)

#define SID_BAD 0xFFFFFFFFu
#define SID_NONE 0xFFFFFFFEu
#define T_INF 0xFFFFFFFFFFFFFFFFull
#define OCC_POS_BITS 40u           // occurrence value = sequence index << 40 | byte offset of the window's last base

struct BuildDev {
    const uint8_t *raw;            // the FASTA bytes
    const uint64_t *seq_off;       // byte offset of every sequence line (ascending)
    const uint32_t *seq_len;       // its length without CR/LF
    uint32_t n_seq;
    uint64_t n_bytes;
    uint32_t lv;                   // complevel
    // label strings
    const char *blob;              // the map's distinct labels, NUL-terminated
    const uint32_t *lab_off;       // per original label
    const uint32_t *seq_label;     // per sequence: original label
    const uint32_t *orig_sid;      // per original label: canonical number of its string
    const uint32_t *cut_sid;       // [label * max_cut + k]: canonical number of the label cut before its k-th ';'
    const uint32_t *sid_rep;       // per canonical string: an original label it is a prefix of ...
    const uint32_t *sid_len;       // ... and its length
    uint32_t max_cut;
};

__device__ __forceinline__ uint32_t code_of(uint8_t c) {           // C2Xb, itree.c:110-121
    switch (c | 0x20) { case 'a': return 0; case 'c': return 1; case 'g': return 2; case 't': return 3; default: return 255; }
}
// Is byte offset e the last base of a k-mer the builder inserts (itree.c:600-621)?  seq: its sequence.
__device__ __forceinline__ bool kmer_at(const BuildDev &d, uint64_t e, uint32_t &seq, uint64_t &word) {
    uint32_t lo = 0, hi = d.n_seq;                                  // largest seq with seq_off <= e
    if (!hi || e < d.seq_off[0]) return false;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (d.seq_off[mid] <= e) lo = mid; else hi = mid; }
    const uint64_t s = d.seq_off[lo], i = e - s;
    const uint32_t kv = 31u + d.lv;
    if (i < kv || i >= d.seq_len[lo]) return false;
    const uint8_t *p = d.raw + s;
    const uint32_t want[4] = {0u, 2u, 1u, 3u};                     // the lv bases before the k-mer: A, G, C, T (itree.c:605-616)
    for (uint32_t c = 0; c < d.lv && c < 4u; ++c) if (code_of(p[i - kv + c]) != want[c]) return false;
    uint64_t w = 0;
    for (uint64_t j = i - 31; j <= i; ++j) {
        const uint32_t c = code_of(p[j]);
        if (c == 255u) return false;                               // itree.c:619: the reference jumps past it; the windows in between fail here too
        w = (w << 2) | c;
    }
    seq = lo; word = w;
    return true;
}
#define B_TILE 4096u               // bytes per block and pass
__global__ void __launch_bounds__(256)
kmer_count_kernel(BuildDev d, uint64_t base, uint64_t n, uint64_t *__restrict__ cnt) {
    __shared__ uint32_t sh[8];
    uint32_t c = 0;
    const uint64_t b0 = base + (uint64_t)blockIdx.x * B_TILE;
    for (uint32_t k = threadIdx.x; k < B_TILE; k += 256) {
        const uint64_t e = b0 + k;
        uint32_t seq; uint64_t w;
        if (e < base + n && kmer_at(d, e, seq, w)) ++c;
    }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < 8; ++w) t += sh[w]; cnt[blockIdx.x] = t; }
}
// writes the block's k-mers in byte order at off[block] (the occurrences come out in input order)
__global__ void __launch_bounds__(256)
kmer_fill_kernel(BuildDev d, uint64_t base, uint64_t n, const uint64_t *__restrict__ off, uint64_t *__restrict__ words, uint64_t *__restrict__ occ) {
    __shared__ uint32_t sh[8];
    __shared__ uint64_t run;
    const uint64_t b0 = base + (uint64_t)blockIdx.x * B_TILE;
    if (threadIdx.x == 0) run = off[blockIdx.x];
    __syncthreads();
    for (uint32_t k0 = 0; k0 < B_TILE; k0 += 256) {
        const uint64_t e = b0 + k0 + threadIdx.x;
        uint32_t seq = 0; uint64_t w = 0;
        const bool ok = e < base + n && kmer_at(d, e, seq, w);
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, ok), lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
        if (lane == 0) sh[wid] = __popc(bal);
        __syncthreads();
        uint64_t at = run;
        for (uint32_t x = 0; x < wid; ++x) at += sh[x];
        if (ok) { at += __popc(bal & ((1u << lane) - 1u)); words[at] = w; occ[at] = ((uint64_t)seq << OCC_POS_BITS) | e; }
        __syncthreads();
        if (threadIdx.x == 0) { uint32_t t = 0; for (int x = 0; x < 8; ++x) t += sh[x]; run += t; }
        __syncthreads();
    }
}
// One thread per distinct word (the first index of its run in the sorted arrays): xeTreeU_RF over its occurrences in
// input order (gg) or xeTreeU (plain build: any second label makes it bad, itree.c:259-265).  fin[i] = canonical
// string number, SID_BAD, or SID_NONE for the other indices of a run.
__global__ void __launch_bounds__(128)
fold_kernel(BuildDev d, const uint64_t *__restrict__ words, const uint64_t *__restrict__ occ, uint64_t n, int gg,
            uint32_t *__restrict__ fin, unsigned long long *__restrict__ first_time, unsigned long long *__restrict__ n_distinct) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t w = words[i];
    if (i && words[i - 1] == w) { fin[i] = SID_NONE; return; }
    atomicAdd(n_distinct, 1ull);
    uint32_t cur = d.orig_sid[d.seq_label[(uint32_t)(occ[i] >> OCC_POS_BITS)]];
    for (uint64_t j = i + 1; j < n && words[j] == w; ++j) {
        if (cur == SID_BAD) break;                                 // already bad (itree.c:286)
        const uint32_t b = d.seq_label[(uint32_t)(occ[j] >> OCC_POS_BITS)];
        if (d.orig_sid[b] == cur) continue;                        // same label (itree.c:285)
        if (!gg) { cur = SID_BAD; break; }
        const uint32_t rep = d.sid_rep[cur], len = d.sid_len[cur];
        const char *olds = d.blob + d.lab_off[rep], *news = d.blob + d.lab_off[b];
        uint32_t num_p = 0;
        for (uint32_t x = 0; x < len && olds[x] == news[x]; ++x) if (olds[x] == ';') ++num_p;   // itree.c:291-294
        if (num_p < 2u) { cur = SID_BAD; break; }                  // critical_cutoff (itree.c:74, 295)
        cur = d.cut_sid[(size_t)rep * d.max_cut + (num_p - 1u)];   // the old label up to its last shared ';' (itree.c:296-299)
        atomicMin(first_time + cur, (unsigned long long)(2ull * (occ[j] & ((1ull << OCC_POS_BITS) - 1ull)) + 1ull));
    }
    fin[i] = cur;
}
__global__ void __launch_bounds__(256)
emit_flag_kernel(const uint32_t *__restrict__ fin, uint64_t n, uint8_t *__restrict__ flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = fin[i] != SID_NONE && fin[i] != SID_BAD;
}
__global__ void __launch_bounds__(256)
emit_ids_kernel(const uint32_t *__restrict__ fin_sel, uint64_t n, const uint32_t *__restrict__ id_of, uint32_t *__restrict__ ids, unsigned long long *__restrict__ counts) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = id_of[fin_sel[i]];
    ids[i] = id;
    atomicAdd(counts + id, 1ull);
}

// the two-column map: name <tab> label (itree.c:506-578, ixCol 0, lblCol 1)
static int read_map(const char *path, std::unordered_map<std::string, std::string> &m, uint64_t *bytes, uint64_t *lines, int *ref_exit) {
    FILE *f = fopen(path, "rb");
    if (!f) { utb_set_error("Invalid input file(s)"); *ref_exit = 1; return UTB_ERR_IO; }
    std::string dump;
    char buf[1 << 16];
    size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) dump.append(buf, k);
    fclose(f);
    *bytes = dump.size();
    if (dump.empty()) { utb_set_error("Input map empty."); *ref_exit = 1; return UTB_ERR_FORMAT; }
    size_t n_lines = std::count(dump.begin(), dump.end(), '\n') + (dump.back() != '\n');
    *lines = n_lines;
    size_t p = 0;
    for (size_t i = 0; i < n_lines; ++i) {
        size_t e = dump.find('\n', p);
        if (e == std::string::npos) e = dump.size();
        const size_t tab = dump.find('\t', p);
        if (tab == std::string::npos || tab >= e) { utb_set_error("Err tab1: %zu", i); *ref_exit = 2; return UTB_ERR_FORMAT; }
        if (tab == p) { utb_set_error("ERROR: map line %zu\nBlank indices are NOT ALLOWED.", i); *ref_exit = 2; return UTB_ERR_FORMAT; }
        size_t le = tab + 1;
        while (le < e && dump[le] != '\r' && dump[le] != '\t') ++le;           // the label ends at a CR or a further tab (itree.c:553)
        if (le == tab + 1) { utb_set_error("ERROR: map line %zu\nBlank labels are NOT ALLOWED.", i + 1); *ref_exit = 2; return UTB_ERR_FORMAT; }
        std::string key = dump.substr(p, tab - p), lab = dump.substr(tab + 1, le - tab - 1);
        if (!m.count(key)) m.emplace(std::move(key), std::move(lab));          // equal keys: the sorted lookup finds one of them; keep the first
        p = e + 1;
    }
    return UTB_OK;
}

extern "C" int utb_build_ubt(const char *fasta_path, const char *map_path, const char *out_path, uint32_t complevel, int gg, int ix_bytes,
                             int device, utb_build_stats *st, int *ref_exit) {
    int dummy;
    if (!ref_exit) ref_exit = &dummy;
    *ref_exit = 0;
    if (!fasta_path || !map_path || !out_path || (ix_bytes != 2 && ix_bytes != 4)) { utb_set_error("utb_build_ubt: bad argument"); return UTB_ERR_ARG; }
    if (complevel > 4) { utb_set_error("complevel %u: the context rule is defined for 0..4 (itree.c:605-616)", complevel); return UTB_ERR_LIMIT; }
    int rc = UTB_OK;
    utb_build_stats S;
    memset(&S, 0, sizeof S);
    // ---- host: map, FASTA framing, label registration ---------------------------------------------------------
    std::unordered_map<std::string, std::string> map;
    int fd = open(fasta_path, O_RDONLY);
    if (fd < 0) { utb_set_error("Invalid input file(s)"); *ref_exit = 1; return UTB_ERR_IO; }
    rc = read_map(map_path, map, &S.map_bytes, &S.map_lines, ref_exit);
    if (rc) { close(fd); return rc; }
    struct stat sb;
    if (fstat(fd, &sb) || sb.st_size <= 0) { close(fd); utb_set_error("Error: no k-mers. Bad input/params!"); *ref_exit = 2; return UTB_ERR_FORMAT; }
    const uint64_t n_bytes = (uint64_t)sb.st_size;
    if (n_bytes >= (1ull << OCC_POS_BITS)) { close(fd); utb_set_error("FASTA larger than 1 TiB"); return UTB_ERR_LIMIT; }
    const char *raw = (const char *)mmap(nullptr, n_bytes, PROT_READ, MAP_PRIVATE, fd, 0);
    if (raw == MAP_FAILED) { close(fd); utb_set_error("mmap failed"); return UTB_ERR_IO; }
    std::vector<uint64_t> seq_off;
    std::vector<uint32_t> seq_len, seq_label;
    std::vector<std::string> labels;                               // original labels in order of first appearance
    std::vector<uint64_t> label_time;
    std::unordered_map<std::string, uint32_t> label_ix;
    {
        uint64_t p = 0, ns = 0;
        while (p < n_bytes) {
            ++ns;
            const char *nl = (const char *)memchr(raw + p, '\n', n_bytes - p);
            const uint64_t he = nl ? (uint64_t)(nl - raw) : n_bytes;
            std::string name(raw + p + 1, he > p ? he - p - 1 : 0);            // the whole header line after its first byte (itree.c:587-590)
            auto it = map.find(name);
            if (it == map.end()) { utb_set_error("Error: taxon map incomplete (line %llu)", (unsigned long long)ns); *ref_exit = 4; rc = UTB_ERR_FORMAT; goto unmap; }
            const uint64_t ss = he + 1;
            if (ss >= n_bytes) { utb_set_error("Error parsing FASTA (1pass): %llu", (unsigned long long)ns); *ref_exit = 2; rc = UTB_ERR_FORMAT; goto unmap; }
            auto li = label_ix.find(it->second);
            uint32_t lab;
            if (li == label_ix.end()) { lab = (uint32_t)labels.size(); label_ix.emplace(it->second, lab); labels.push_back(it->second); label_time.push_back(2ull * ss); }
            else lab = li->second;
            const char *nl2 = (const char *)memchr(raw + ss, '\n', n_bytes - ss);
            uint64_t se = nl2 ? (uint64_t)(nl2 - raw) : n_bytes, len = se - ss;
            if (len && raw[ss + len - 1] == '\r') --len;                        // itree.c:598-599
            if (len >= 0xFFFFFFFFull) { utb_set_error("sequence line too long"); rc = UTB_ERR_LIMIT; goto unmap; }
            seq_off.push_back(ss); seq_len.push_back((uint32_t)len); seq_label.push_back(lab);
            p = se + 1;
        }
    }
    {
        // ---- every string that can become a label: the labels and their cuts before a ';' ---------------------------
        const uint32_t n_lab = (uint32_t)labels.size();
        std::unordered_map<std::string, uint32_t> sid_of;
        std::vector<uint32_t> sid_rep, sid_len, orig_sid(n_lab), lab_off(n_lab);
        std::string blob;
        uint32_t max_cut = 1;
        for (uint32_t a = 0; a < n_lab; ++a) max_cut = std::max<uint32_t>(max_cut, (uint32_t)std::count(labels[a].begin(), labels[a].end(), ';'));
        std::vector<uint32_t> cut_sid((size_t)n_lab * max_cut, SID_BAD);
        auto sid_for = [&](const std::string &s, uint32_t rep, uint32_t len) {
            auto f = sid_of.find(s);
            if (f != sid_of.end()) return f->second;
            const uint32_t id = (uint32_t)sid_rep.size();
            sid_of.emplace(s, id); sid_rep.push_back(rep); sid_len.push_back(len);
            return id;
        };
        for (uint32_t a = 0; a < n_lab; ++a) {
            lab_off[a] = (uint32_t)blob.size();
            blob += labels[a]; blob.push_back('\0');
            orig_sid[a] = sid_for(labels[a], a, (uint32_t)labels[a].size());
            uint32_t k = 0;
            for (uint32_t x = 0; x < labels[a].size(); ++x) if (labels[a][x] == ';') cut_sid[(size_t)a * max_cut + k++] = sid_for(labels[a].substr(0, x), a, x);
        }
        const uint32_t n_sid = (uint32_t)sid_rep.size();
        std::vector<unsigned long long> first_time(n_sid, T_INF);
        for (uint32_t a = 0; a < n_lab; ++a) first_time[orig_sid[a]] = std::min<unsigned long long>(first_time[orig_sid[a]], label_time[a]);
        const uint32_t n_seq = (uint32_t)seq_off.size();
        S.sequences = n_seq;

        // ---- device ------------------------------------------------------------------------------------------------------
        uint8_t *d_raw = nullptr, *d_flag = nullptr; uint64_t *d_seq_off = nullptr, *d_words = nullptr, *d_occ = nullptr, *d_words2 = nullptr, *d_occ2 = nullptr, *d_boff = nullptr;
        uint32_t *d_seq_len = nullptr, *d_seq_label = nullptr, *d_lab_off = nullptr, *d_orig_sid = nullptr, *d_cut_sid = nullptr, *d_sid_rep = nullptr, *d_sid_len = nullptr;
        uint64_t *d_cnt = nullptr; uint32_t *d_fin = nullptr, *d_fin_sel = nullptr, *d_id_of = nullptr, *d_ids = nullptr;
        uint64_t *d_words_sel = nullptr; unsigned long long *d_first = nullptr, *d_misc = nullptr, *d_counts = nullptr;
        char *d_blob = nullptr; void *d_tmp = nullptr; size_t tmp_bytes = 0;
        std::vector<uint32_t> id_of; std::vector<uint32_t> order; std::vector<unsigned long long> counts;
        std::vector<uint64_t> h_words; std::vector<uint32_t> h_ids;
        uint64_t n_occ = 0, n_good = 0;
        unsigned long long h_misc[2] = {0, 0};
        const uint32_t n_blocks = (uint32_t)((n_bytes + B_TILE - 1) / B_TILE);
        BuildDev D;
        cudaError_t e0 = cudaSetDevice(device);
        if (e0 != cudaSuccess) { utb_set_error("no usable CUDA device (%s); there is no CPU fallback", cudaGetErrorString(e0)); rc = UTB_ERR_CUDA; goto done; }
        BCK(cudaMalloc(&d_raw, n_bytes + 64));
        BCK(cudaMemcpy(d_raw, raw, n_bytes, cudaMemcpyHostToDevice));
        BCK(cudaMalloc(&d_seq_off, (n_seq + 1) * 8ull)); BCK(cudaMalloc(&d_seq_len, (n_seq + 1) * 4ull)); BCK(cudaMalloc(&d_seq_label, (n_seq + 1) * 4ull));
        BCK(cudaMemcpy(d_seq_off, seq_off.data(), n_seq * 8ull, cudaMemcpyHostToDevice));
        BCK(cudaMemcpy(d_seq_len, seq_len.data(), n_seq * 4ull, cudaMemcpyHostToDevice));
        BCK(cudaMemcpy(d_seq_label, seq_label.data(), n_seq * 4ull, cudaMemcpyHostToDevice));
        BCK(cudaMalloc(&d_blob, blob.size() + 64)); BCK(cudaMemcpy(d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
        BCK(cudaMalloc(&d_lab_off, n_lab * 4ull)); BCK(cudaMemcpy(d_lab_off, lab_off.data(), n_lab * 4ull, cudaMemcpyHostToDevice));
        BCK(cudaMalloc(&d_orig_sid, n_lab * 4ull)); BCK(cudaMemcpy(d_orig_sid, orig_sid.data(), n_lab * 4ull, cudaMemcpyHostToDevice));
        BCK(cudaMalloc(&d_cut_sid, cut_sid.size() * 4ull)); BCK(cudaMemcpy(d_cut_sid, cut_sid.data(), cut_sid.size() * 4ull, cudaMemcpyHostToDevice));
        BCK(cudaMalloc(&d_sid_rep, n_sid * 4ull)); BCK(cudaMemcpy(d_sid_rep, sid_rep.data(), n_sid * 4ull, cudaMemcpyHostToDevice));
        BCK(cudaMalloc(&d_sid_len, n_sid * 4ull)); BCK(cudaMemcpy(d_sid_len, sid_len.data(), n_sid * 4ull, cudaMemcpyHostToDevice));
        BCK(cudaMalloc(&d_first, n_sid * 8ull)); BCK(cudaMemcpy(d_first, first_time.data(), n_sid * 8ull, cudaMemcpyHostToDevice));
        BCK(cudaMalloc(&d_misc, 16)); BCK(cudaMemset(d_misc, 0, 16));
        D.raw = d_raw; D.seq_off = d_seq_off; D.seq_len = d_seq_len; D.n_seq = n_seq; D.n_bytes = n_bytes; D.lv = complevel;
        D.blob = d_blob; D.lab_off = d_lab_off; D.seq_label = d_seq_label; D.orig_sid = d_orig_sid; D.cut_sid = d_cut_sid;
        D.sid_rep = d_sid_rep; D.sid_len = d_sid_len; D.max_cut = max_cut;
        // 1. k-mers in input order: count per 4 KiB block, scan, fill
        BCK(cudaMalloc(&d_cnt, (n_blocks + 1) * 8ull)); BCK(cudaMalloc(&d_boff, (n_blocks + 1) * 8ull));
        kmer_count_kernel<<<n_blocks, 256>>>(D, 0, n_bytes, d_cnt);
        BCK(cudaGetLastError());
        BCK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cnt, d_boff, (int)n_blocks + 1));
        BCK(cudaMalloc(&d_tmp, tmp_bytes));
        BCK(cudaMemset(d_cnt + n_blocks, 0, 8));
        BCK(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_cnt, d_boff, (int)n_blocks + 1));
        BCK(cudaMemcpy(&n_occ, d_boff + n_blocks, 8, cudaMemcpyDeviceToHost));
        cudaFree(d_tmp); d_tmp = nullptr;
        S.kmers_seen = n_occ;
        if (!n_occ) { utb_set_error("Error: no k-mers. Bad input/params!"); *ref_exit = 2; rc = UTB_ERR_FORMAT; goto done; }   // itree.c:630
        if (n_occ >= (1ull << 31)) { utb_set_error("%llu k-mer occurrences: more than one pass of this builder holds", (unsigned long long)n_occ); rc = UTB_ERR_LIMIT; goto done; }
        BCK(cudaMalloc(&d_words, n_occ * 8)); BCK(cudaMalloc(&d_occ, n_occ * 8)); BCK(cudaMalloc(&d_words2, n_occ * 8)); BCK(cudaMalloc(&d_occ2, n_occ * 8));
        kmer_fill_kernel<<<n_blocks, 256>>>(D, 0, n_bytes, d_boff, d_words, d_occ);
        BCK(cudaGetLastError());
        // 2. stable sort by word: the occurrences of a word stay in input order
        BCK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_words, d_words2, d_occ, d_occ2, n_occ));
        BCK(cudaMalloc(&d_tmp, tmp_bytes));
        BCK(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_words, d_words2, d_occ, d_occ2, n_occ));
        cudaFree(d_tmp); d_tmp = nullptr;
        cudaFree(d_words); d_words = nullptr; cudaFree(d_occ); d_occ = nullptr;
        // 3. the fold
        BCK(cudaMalloc(&d_fin, n_occ * 4));
        fold_kernel<<<(unsigned)((n_occ + 127) / 128), 128>>>(D, d_words2, d_occ2, n_occ, gg, d_fin, d_first, d_misc);
        BCK(cudaGetLastError());
        BCK(cudaMemcpy(h_misc, d_misc, 16, cudaMemcpyDeviceToHost));
        S.kmers_made = h_misc[0];                                  // distinct words, the bad ones included ("k-mers made", itree.c:629)
        BCK(cudaMemcpy(first_time.data(), d_first, n_sid * 8ull, cudaMemcpyDeviceToHost));
        // 4. ids: every string ever registered, in order of registration
        for (uint32_t s = 0; s < n_sid; ++s) if (first_time[s] != T_INF) order.push_back(s);
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return first_time[a] < first_time[b]; });
        if (order.size() > (ix_bytes == 2 ? 65534u : 0xFFFFFFFDu)) { utb_set_error("%zu labels do not fit a %d-byte IXTYPE", order.size(), ix_bytes); rc = UTB_ERR_LIMIT; goto done; }
        id_of.assign(n_sid, SID_BAD);
        for (uint32_t k = 0; k < order.size(); ++k) id_of[order[k]] = k;
        S.labels = (uint32_t)order.size();
        // 5. good words in order + their ids + per-label counts
        BCK(cudaMalloc(&d_flag, n_occ));
        emit_flag_kernel<<<(unsigned)((n_occ + 255) / 256), 256>>>(d_fin, n_occ, d_flag);
        BCK(cudaMalloc(&d_words_sel, n_occ * 8)); BCK(cudaMalloc(&d_fin_sel, n_occ * 4));
        BCK(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, d_words2, d_flag, d_words_sel, d_misc + 1, (int)n_occ));
        BCK(cudaMalloc(&d_tmp, tmp_bytes));
        BCK(cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, d_words2, d_flag, d_words_sel, d_misc + 1, (int)n_occ));
        BCK(cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, d_fin, d_flag, d_fin_sel, d_misc + 1, (int)n_occ));
        BCK(cudaMemcpy(h_misc, d_misc, 16, cudaMemcpyDeviceToHost));
        n_good = h_misc[1];
        S.records = n_good;
        BCK(cudaMalloc(&d_id_of, n_sid * 4ull)); BCK(cudaMemcpy(d_id_of, id_of.data(), n_sid * 4ull, cudaMemcpyHostToDevice));
        BCK(cudaMalloc(&d_counts, (order.size() + 1) * 8ull)); BCK(cudaMemset(d_counts, 0, (order.size() + 1) * 8ull));
        BCK(cudaMalloc(&d_ids, (n_good + 1) * 4ull));
        if (n_good) emit_ids_kernel<<<(unsigned)((n_good + 255) / 256), 256>>>(d_fin_sel, n_good, d_id_of, d_ids, d_counts);
        BCK(cudaGetLastError());
        h_words.resize(n_good); h_ids.resize(n_good); counts.resize(order.size());
        if (n_good) { BCK(cudaMemcpy(h_words.data(), d_words_sel, n_good * 8, cudaMemcpyDeviceToHost)); BCK(cudaMemcpy(h_ids.data(), d_ids, n_good * 4, cudaMemcpyDeviceToHost)); }
        BCK(cudaMemcpy(counts.data(), d_counts, order.size() * 8ull, cudaMemcpyDeviceToHost));
        // ---- the .ubt (itree.c:1317-1343) and the .log (itree.c:1225-1232) ------------------------------------------------
        {
            FILE *of = fopen(out_path, "wb");
            if (!of) { utb_set_error("Invalid output filename"); rc = UTB_ERR_IO; goto done; }
            setvbuf(of, nullptr, _IOFBF, (size_t)8 << 20);
            const uint64_t md[4] = {8, 0, (uint64_t)ix_bytes, n_good};
            fwrite(md, 8, 4, of);
            const size_t rs = 8 + (size_t)ix_bytes, BLK = 1 << 20;
            std::vector<uint8_t> buf(BLK * rs);
            for (uint64_t i0 = 0; i0 < n_good; i0 += BLK) {
                const size_t c = (size_t)std::min<uint64_t>(BLK, n_good - i0);
                for (size_t j = 0; j < c; ++j) { memcpy(&buf[j * rs], &h_words[i0 + j], 8); memcpy(&buf[j * rs + 8], &h_ids[i0 + j], ix_bytes); }
                if (fwrite(buf.data(), rs, c, of) != c) { fclose(of); utb_set_error("write error on %s", out_path); rc = UTB_ERR_IO; goto done; }
            }
            std::string tail;
            for (uint32_t k = 0; k < order.size(); ++k) {
                const uint32_t s = order[k];
                tail.append(blob.data() + lab_off[sid_rep[s]], sid_len[s]);
                tail += '\t'; tail += std::to_string(counts[k]); tail += '\n';
            }
            fwrite(tail.data(), 1, tail.size(), of);
            if (fclose(of)) { utb_set_error("write error on %s", out_path); rc = UTB_ERR_IO; goto done; }
            std::string log_path = std::string(out_path) + (gg ? ".gg" : "") + ".log";
            FILE *lf = fopen(log_path.c_str(), "wb");
            if (lf) { fwrite(tail.data(), 1, tail.size(), lf); fclose(lf); }
        }
    done:
        cudaFree(d_raw); cudaFree(d_flag); cudaFree(d_seq_off); cudaFree(d_words); cudaFree(d_occ); cudaFree(d_words2); cudaFree(d_occ2); cudaFree(d_boff);
        cudaFree(d_seq_len); cudaFree(d_seq_label); cudaFree(d_lab_off); cudaFree(d_orig_sid); cudaFree(d_cut_sid); cudaFree(d_sid_rep); cudaFree(d_sid_len);
        cudaFree(d_cnt); cudaFree(d_fin); cudaFree(d_fin_sel); cudaFree(d_id_of); cudaFree(d_ids); cudaFree(d_words_sel); cudaFree(d_first); cudaFree(d_misc);
        cudaFree(d_counts); cudaFree(d_blob); cudaFree(d_tmp);
    }
unmap:
    munmap((void *)raw, n_bytes);
    close(fd);
    if (st) *st = S;
    return rc;
}

// The BUILD_GG binary's CLI contract (itree.c:1379-1408): utree-build_gg input_fasta.fa labels.map output.ubt threads [complevel]
extern "C" int utb_build_main(int argc, char **argv) {
    if (argc < 5) { printf("[v2.0RF SigNature Edition] usage: utree-buildGG input_fasta.fa labels.map output.ubt threads{0=auto} [complevel]\n"); return 1; }
    printf("This is UTree [v2.0RF SigNature Edition]\n");
    long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    int threads = atoi(argv[4]) ? atoi(argv[4]) : (int)(ncpu > 0 ? ncpu : 1);
    printf("Using up to %d threads.\n", threads);
    puts("Tree initialized.");
    uint32_t cl = argc > 5 ? (uint32_t)atoi(argv[5]) : 1u;
    printf("Setting compression level to %u\n", cl);
    const char *ixe = getenv("UTB_IX_BYTES");                      // the reference is compiled per IXTYPE; here one binary, default uint16_t
    utb_build_stats st;
    int ref_exit = 0;
    int rc = utb_build_ubt(argv[1], argv[2], argv[3], cl, 1, ixe && atoi(ixe) == 4 ? 4 : 2, 0, &st, &ref_exit);
    if (st.map_lines && !(rc && ref_exit == 1)) printf("Parsed map. %llu bytes, %llu lines.\n", (unsigned long long)st.map_bytes, (unsigned long long)st.map_lines);
    if (rc) {
        if (rc == UTB_ERR_CUDA) { fprintf(stderr, "utree-b200: %s\n", utb_last_error()); return 6; }   // no GPU: there is no CPU fallback
        puts(utb_last_error());
        return ref_exit ? ref_exit : 6;
    }
    printf("Done with sequence parse: %llu k-mers made\n", (unsigned long long)st.kmers_made);
    puts("File parsed.");
    unsigned long long total = st.records;
    printf("Total nodes in tree: %llu [%llu labels]\n", total, (unsigned long long)st.labels);
    puts("Tree written.");
    return 0;
}
