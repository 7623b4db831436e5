// GPU FASTA ingest and background residue histogram (N3).
//
// Raw FASTA text -> 1-byte residue codes + int64 offsets, with the reference reader's exact semantics
// (fastareader, plaac.java:4302-4375; string2aa :1764-1769; terminal '*' strip :758):
//   * lines end at \n, \r or \r\n (BufferedReader.readLine);
//   * a line whose first byte is '>' starts a record; its name is the rest of that line;
//   * the record's sequence is every byte (NOT trimmed: blanks, digits ... become code 0 = X) of the following
//     lines up to the next '>' line, an EMPTY line, or the end of the file;
//   * after an empty line everything up to the next '>' line is skipped; so is everything before the first '>';
//   * one trailing '*' of the whole sequence is dropped.
// A byte is a residue iff the latest "header or empty" line start at or before it is a header that is not its own
// line.  That is two running maxima and one running count over the bytes, i.e. a prefix scan:
//
//   k_ing_reduce   per 4 KB tile: (last line start, last header/empty line start, #headers)
//   k_ing_scan3    exclusive scan of the tile aggregates (one CTA)
//   k_ing_apply<0> per tile: residue flags from the scanned state -> residues per tile
//   k_scan_exclusive (prep.cuh) over those counts
//   k_ing_apply<1> per tile: the same flags again, local ranks, scatter of codes / offsets / name spans
//
// The text is read three times and the codes written once: HBM-bound byte work, no tensor cores.
// k_bg_hist: computeaafreq/countaas/isvalidprotein (:1655-1739) on the ingested records, 64-bit counters.
#pragma once
#include "common.cuh"

namespace plaac {

constexpr int kIngThreads = 256;
constexpr int kIngPer = 16;
constexpr int kIngTile = kIngThreads * kIngPer;

struct Ing3 {
    long long ls;  // position of the last line start (max), -1 = none
    long long mk;  // (position << 1 | is_header) of the last header-or-empty line start (max), -1 = none
    long long nh;  // number of header lines (sum)
};
__device__ __forceinline__ Ing3 ing_combine(const Ing3& a, const Ing3& b)
{
    Ing3 r;
    r.ls = a.ls > b.ls ? a.ls : b.ls;
    r.mk = a.mk > b.mk ? a.mk : b.mk;
    r.nh = a.nh + b.nh;
    return r;
}
__device__ __forceinline__ Ing3 ing_identity()
{
    Ing3 r;
    r.ls = -1;
    r.mk = -1;
    r.nh = 0;
    return r;
}
__device__ __forceinline__ bool is_term(unsigned char c) { return c == '\n' || c == '\r'; }

// per-byte events: advances the running state st over byte i
__device__ __forceinline__ void ing_step(Ing3& st, long long i, unsigned char c, unsigned char prev)
{
    const bool ls = (i == 0) || prev == '\n' || (prev == '\r' && c != '\n');
    if (ls) {
        st.ls = i;
        if (is_term(c))
            st.mk = i << 1;  // empty line
        else if (c == '>') {
            st.mk = (i << 1) | 1;
            st.nh += 1;
        }
    }
}

// aatoint, plaac.java:1508-1534 (case-insensitive; '*' -> 21; everything else -> 0)
__device__ __forceinline__ uint8_t aa_code(unsigned char c)
{
    const unsigned char u = c & 0xdf;  // fold case for letters
    switch (u) {
        case 'A': return 1; case 'C': return 2; case 'D': return 3; case 'E': return 4; case 'F': return 5;
        case 'G': return 6; case 'H': return 7; case 'I': return 8; case 'K': return 9; case 'L': return 10;
        case 'M': return 11; case 'N': return 12; case 'P': return 13; case 'Q': return 14; case 'R': return 15;
        case 'S': return 16; case 'T': return 17; case 'V': return 18; case 'W': return 19; case 'Y': return 20;
        default: break;
    }
    // only real letters fold: bytes such as '!' (0x21) & 0xdf = 0x01 never reach a case above; '*' is exact
    return c == '*' ? 21 : 0;
}

struct IngTileLoad {
    unsigned char b[kIngPer];
    unsigned char prev;
    long long i0;
    int nvalid;
};

__device__ __forceinline__ IngTileLoad ing_load(const unsigned char* __restrict__ text, long long nbytes)
{
    IngTileLoad L;
    L.i0 = ((long long)blockIdx.x * kIngThreads + threadIdx.x) * kIngPer;
    L.nvalid = (int)max(0LL, min((long long)kIngPer, nbytes - L.i0));
    if (L.nvalid == kIngPer && (((uintptr_t)(text + L.i0)) & 15) == 0) {
        const uint4 v = *reinterpret_cast<const uint4*>(text + L.i0);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < kIngPer; k++) L.b[k] = (unsigned char)(w[k >> 2] >> ((k & 3) * 8));
    } else {
#pragma unroll
        for (int k = 0; k < kIngPer; k++) L.b[k] = k < L.nvalid ? text[L.i0 + k] : (unsigned char)'\n';
    }
    L.prev = (L.i0 > 0 && L.i0 <= nbytes) ? text[L.i0 - 1] : (unsigned char)'\n';
    return L;
}

// exclusive scan of one Ing3 per thread across the CTA; returns the exclusive prefix, *total = CTA aggregate
__device__ __forceinline__ Ing3 ing_block_exscan(const Ing3& mine, Ing3* total)
{
    __shared__ Ing3 warp_tot[kIngThreads / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    Ing3 inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Ing3 o;
        o.ls = __shfl_up_sync(0xffffffffu, inc.ls, d);
        o.mk = __shfl_up_sync(0xffffffffu, inc.mk, d);
        o.nh = __shfl_up_sync(0xffffffffu, inc.nh, d);
        if (lane >= d) inc = ing_combine(o, inc);
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    Ing3 ex;
    ex.ls = __shfl_up_sync(0xffffffffu, inc.ls, 1);
    ex.mk = __shfl_up_sync(0xffffffffu, inc.mk, 1);
    ex.nh = __shfl_up_sync(0xffffffffu, inc.nh, 1);
    if (lane == 0) ex = ing_identity();
    Ing3 pre = ing_identity();
    for (int w = 0; w < wid; w++) pre = ing_combine(pre, warp_tot[w]);
    if (total) {
        Ing3 t = ing_identity();
        for (int w = 0; w < kIngThreads / 32; w++) t = ing_combine(t, warp_tot[w]);
        *total = t;
    }
    __syncthreads();
    return ing_combine(pre, ex);
}

__global__ void __launch_bounds__(kIngThreads)
k_ing_reduce(const unsigned char* __restrict__ text, long long nbytes, Ing3* __restrict__ agg)
{
    const IngTileLoad L = ing_load(text, nbytes);
    Ing3 st = ing_identity();
    unsigned char prev = L.prev;
#pragma unroll
    for (int k = 0; k < kIngPer; k++) {
        if (k < L.nvalid) ing_step(st, L.i0 + k, L.b[k], prev);
        prev = L.b[k];
    }
    Ing3 total;
    ing_block_exscan(st, &total);
    if (threadIdx.x == 0) agg[blockIdx.x] = total;
}

// exclusive scan of the tile aggregates, one CTA of 1024 threads; agg is overwritten by the exclusive prefixes and
// *grand receives the total
__global__ void __launch_bounds__(1024) k_ing_scan3(Ing3* __restrict__ agg, long long ntiles, Ing3* __restrict__ grand)
{
    __shared__ Ing3 warp_tot[32];
    __shared__ Ing3 carry_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = ing_identity();
    __syncthreads();
    for (long long base = 0; base < ntiles; base += 1024) {
        const long long i = base + threadIdx.x;
        const Ing3 mine = i < ntiles ? agg[i] : ing_identity();
        Ing3 inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            Ing3 o;
            o.ls = __shfl_up_sync(0xffffffffu, inc.ls, d);
            o.mk = __shfl_up_sync(0xffffffffu, inc.mk, d);
            o.nh = __shfl_up_sync(0xffffffffu, inc.nh, d);
            if (lane >= d) inc = ing_combine(o, inc);
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        Ing3 ex;
        ex.ls = __shfl_up_sync(0xffffffffu, inc.ls, 1);
        ex.mk = __shfl_up_sync(0xffffffffu, inc.mk, 1);
        ex.nh = __shfl_up_sync(0xffffffffu, inc.nh, 1);
        if (lane == 0) ex = ing_identity();
        Ing3 pre = carry_s;
        for (int w = 0; w < wid; w++) pre = ing_combine(pre, warp_tot[w]);
        if (i < ntiles) agg[i] = ing_combine(pre, ex);
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = ing_combine(ing_combine(pre, ex), mine);
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand = carry_s;
}

struct IngOut {
    uint8_t* codes;        // residue codes, records concatenated
    long long* offsets;    // nrec + 1
    long long* name_pos;   // first byte of the record name in the text (after '>')
    int32_t* name_len;     // bytes up to the line end
    uint8_t* flags;        // bit 0: name found after an empty line / at file start (the jar trims it, :4362);
                           // bit 1: a terminal '*' was stripped (zeroed before the launch)
    long long max_rec;     // capacity of the per-record arrays (records beyond it are counted, not stored)
};

template <int SCATTER>
__global__ void __launch_bounds__(kIngThreads)
k_ing_apply(const unsigned char* __restrict__ text, long long nbytes, const Ing3* __restrict__ tile_pre,
            int32_t* __restrict__ tile_cnt, const long long* __restrict__ tile_base, IngOut out)
{
    __shared__ int warp_cnt[kIngThreads / 32];
    const IngTileLoad L = ing_load(text, nbytes);
    // thread aggregate, then the state entering this thread's first byte
    Ing3 st = ing_identity();
    {
        unsigned char prev = L.prev;
#pragma unroll
        for (int k = 0; k < kIngPer; k++) {
            if (k < L.nvalid) ing_step(st, L.i0 + k, L.b[k], prev);
            prev = L.b[k];
        }
    }
    const Ing3 ex = ing_combine(tile_pre[blockIdx.x], ing_block_exscan(st, nullptr));
    // second sweep with the true incoming state: residue flags
    Ing3 run = ex;
    uint32_t dmask = 0, hmask = 0, starmask = 0;
    unsigned char prev = L.prev;
#pragma unroll
    for (int k = 0; k < kIngPer; k++) {
        if (k < L.nvalid) {
            const long long i = L.i0 + k;
            const unsigned char c = L.b[k];
            const long long mk_before = run.mk;
            ing_step(run, i, c, prev);
            const bool in_rec = run.mk >= 0 && (run.mk & 1);
            const bool header_line = in_rec && (run.mk >> 1) == run.ls;
            if (header_line && run.ls == i) {
                hmask |= 1u << k;
                if (!(mk_before >= 0 && (mk_before & 1))) hmask |= 1u << (16 + k);  // found by hasmorefastas
            }
            bool d = in_rec && !header_line && !is_term(c);
            if (d && c == '*') {
                // terminal iff nothing but one line terminator, then an empty line / a header / the end follows
                long long j = i + 1;
                bool terminal;
                if (j >= nbytes)
                    terminal = true;
                else {
                    const unsigned char c1 = text[j];
                    if (!is_term(c1))
                        terminal = false;
                    else {
                        j++;
                        if (c1 == '\r' && j < nbytes && text[j] == '\n') j++;
                        terminal = j >= nbytes || is_term(text[j]) || text[j] == '>';
                    }
                }
                if (terminal) {
                    d = false;
                    starmask |= 1u << k;
                }
            }
            if (d) dmask |= 1u << k;
        }
        prev = L.b[k];
    }
    // exclusive rank of this thread's residues inside the tile
    const int mine = __popc(dmask);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) warp_cnt[wid] = inc;
    __syncthreads();
    int pre = 0, tot = 0;
    for (int w = 0; w < kIngThreads / 32; w++) {
        if (w < wid) pre += warp_cnt[w];
        tot += warp_cnt[w];
    }
    if (!SCATTER) {
        if (threadIdx.x == 0) tile_cnt[blockIdx.x] = tot;
        return;
    }
    long long dst = tile_base[blockIdx.x] + pre + (inc - mine);
    long long nh = ex.nh;
#pragma unroll
    for (int k = 0; k < kIngPer; k++) {
        if (hmask & (1u << k)) {
            const long long r = nh++;
            if (r < out.max_rec) {
                const long long i = L.i0 + k;
                out.offsets[r] = dst;
                out.name_pos[r] = i + 1;
                long long e = i + 1;
                while (e < nbytes && !is_term(text[e])) e++;
                out.name_len[r] = (int32_t)(e - (i + 1));
                if (hmask & (1u << (16 + k)))  // flags[] was zeroed; word-wide atomics because neighbours share the word
                    atomicOr(reinterpret_cast<unsigned int*>(out.flags + (r & ~3LL)), 1u << (8 * (int)(r & 3)));
            }
        }
        if (starmask & (1u << k)) {
            // the star belongs to the record whose header precedes it
            const long long r = nh - 1;
            if (r >= 0 && r < out.max_rec) {
                unsigned int* wptr = reinterpret_cast<unsigned int*>(out.flags + (r & ~3LL));
                atomicOr(wptr, 2u << (8 * (int)(r & 3)));
            }
        }
        if (dmask & (1u << k)) out.codes[dst++] = aa_code(L.b[k]);
    }
}

// computeaafreq :1655-1666: 22-bin histogram over the UNSTRIPPED sequences of the valid records
// (isvalidprotein :1732-1739: no X or '*' at positions 1..m-2, no X at the last position; position 0 is never
// checked).  One warp per record; counters are 64-bit (the jar's ints overflow above 2^31 per bin).
__global__ void __launch_bounds__(256)
k_bg_hist(const uint8_t* __restrict__ codes, const long long* __restrict__ offsets, const uint8_t* __restrict__ flags,
          long long nrec, unsigned long long* __restrict__ hist)
{
    __shared__ unsigned int sh[8][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    sh[wid][lane] = 0;
    __syncwarp();
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    unsigned int flushed = 0;
    for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nrec; r += warps) {
        const long long lo = offsets[r];
        const int n = (int)(offsets[r + 1] - lo);
        const bool star = (flags[r] & 2) != 0;
        const int m = n + (star ? 1 : 0);  // unstripped length
        if (m < 1) continue;               // the jar would throw on an empty record; nothing to count
        bool bad = false;
        for (int i = lane; i < n; i += 32) {
            const uint8_t c = codes[lo + i];
            const bool interior = i >= 1 && i <= m - 2;
            const bool last = i == m - 1;
            if (interior && (c == 0 || c == 21)) bad = true;
            if (last && c == 0) bad = true;
        }
        if (__any_sync(0xffffffffu, bad)) continue;
        for (int i = lane; i < n; i += 32) atomicAdd(&sh[wid][codes[lo + i] & 31], 1u);
        if (star && lane == 0) atomicAdd(&sh[wid][21], 1u);
        flushed += (unsigned)n + 1;
        if (flushed > 0x40000000u) {  // keep the 32-bit shared counters far from overflow
            __syncwarp();
            const unsigned int v = sh[wid][lane];
            if (v) atomicAdd(&hist[lane], (unsigned long long)v);
            sh[wid][lane] = 0;
            flushed = 0;
            __syncwarp();
        }
    }
    __syncwarp();
    const unsigned int v = sh[wid][lane];
    if (v && lane < PLAAC_NAA) atomicAdd(&hist[lane], (unsigned long long)v);
}

}  // namespace plaac
