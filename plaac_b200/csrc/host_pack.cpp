// Host side of the transport-lean batch call (include/plaac_cuda.h): residue codes <-> radix-22 words, 7 residues per
// uint32 (22^7 = 2 494 357 888 < 2^32).  Pure host code; the device side is k_unpack22 (lean.cuh).
// The encoding of a residue is aatoint's (plaac.java:1508-1534); the packing itself has no reference counterpart.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/plaac_cuda.h"

namespace {

constexpr int kPer = PLAAC_PACK_PER_WORD;

struct CharLut {
    uint8_t t[256];
    CharLut()
    {
        static const char names[] = "XACDEFGHIKLMNPQRSTVWY";
        memset(t, 0, sizeof(t));
        for (int i = 1; i <= 20; i++) {
            t[(unsigned char)names[i]] = (uint8_t)i;
            t[(unsigned char)(names[i] + 32)] = (uint8_t)i;
        }
        t[(unsigned char)'*'] = 21;
    }
};

// words [w0, w1) from the symbols src[7*w ..]; MAP: symbol -> code.  Returns 1 if a symbol was out of range.
template <class MAP>
int pack_range(const uint8_t* src, int64_t nres, uint32_t* words, int64_t w0, int64_t w1, MAP map)
{
    int bad = 0;
    for (int64_t w = w0; w < w1; w++) {
        const int64_t r0 = w * kPer;
        const int m = (int)std::min<int64_t>(kPer, nres - r0);
        uint32_t v = 0;
        for (int k = m - 1; k >= 0; k--) {
            uint32_t c = map(src[r0 + k]);
            if (c > 21u) {
                c = 0;
                bad = 1;
            }
            v = v * 22u + c;
        }
        words[w] = v;
    }
    return bad;
}

template <class MAP>
int pack_threads(const uint8_t* src, int64_t nres, uint32_t* words, int nthreads, MAP map)
{
    if (nres < 0 || (nres > 0 && (!src || !words))) return PLAAC_E_INVALID;
    const int64_t nw = (nres + kPer - 1) / kPer;
    unsigned hw = std::thread::hardware_concurrency();
    int64_t nt = nthreads > 0 ? nthreads : (hw ? hw : 1);
    nt = std::max<int64_t>(1, std::min<int64_t>(nt, nw / (1 << 18)));  // at least 256 k words (1 MB) per thread
    std::vector<int> bad((size_t)nt, 0);
    std::vector<std::thread> th;
    const int64_t per = (nw + nt - 1) / nt;
    for (int64_t t = 1; t < nt; t++) {
        const int64_t lo = std::min(nw, t * per), hi = std::min(nw, lo + per);
        if (hi <= lo) continue;
        int* b = &bad[(size_t)t];
        try {
            th.emplace_back([=] { *b = pack_range(src, nres, words, lo, hi, map); });
        } catch (...) {  // no thread to be had: this slice on the calling thread
            *b = pack_range(src, nres, words, lo, hi, map);
        }
    }
    bad[0] = pack_range(src, nres, words, 0, std::min(nw, per), map);
    for (auto& t : th) t.join();
    for (int b : bad)
        if (b) return PLAAC_E_INVALID;
    return PLAAC_OK;
}

}  // namespace

extern "C" {

int64_t plaac_packed_words(int64_t nres) { return nres <= 0 ? 0 : (nres + kPer - 1) / kPer; }

int plaac_pack_host(const uint8_t* codes, int64_t nres, uint32_t* words, int nthreads)
try {
    return pack_threads(codes, nres, words, nthreads, [](uint8_t c) -> uint32_t { return c; });
} catch (...) {
    return PLAAC_E_NOMEM;
}

int plaac_pack_chars_host(const char* chars, int64_t nres, uint32_t* words, int nthreads)
try {
    static const CharLut lut;
    const uint8_t* t = lut.t;
    return pack_threads(reinterpret_cast<const uint8_t*>(chars), nres, words, nthreads,
                        [t](uint8_t c) -> uint32_t { return t[c]; });
} catch (...) {
    return PLAAC_E_NOMEM;
}

int plaac_pack_append_host(const uint8_t* codes, int64_t nres, uint32_t* words, int64_t pos, int nthreads)
try {
    if (nres < 0 || pos < 0 || (nres > 0 && (!codes || !words))) return PLAAC_E_INVALID;
    int bad = 0;
    int64_t done = 0;
    const int k0 = (int)(pos % kPer);
    if (k0 != 0 && nres > 0) {
        // the first word already holds k0 digits of earlier residues: add ours above them
        const int64_t w = pos / kPer;
        uint32_t scale = 1;
        for (int i = 0; i < k0; i++) scale *= 22u;
        uint32_t v = words[w] % scale;  // (digits above k0 are rewritten)
        for (int k = k0; k < kPer && done < nres; k++, done++) {
            uint32_t c = codes[done];
            if (c > 21u) c = 0, bad = 1;
            v += c * scale;
            scale *= 22u;
        }
        words[w] = v;
    }
    if (done < nres) {
        const int rc = pack_threads(codes + done, nres - done, words + (pos + done) / kPer, nthreads, [](uint8_t c) -> uint32_t { return c; });
        if (rc != PLAAC_OK) return rc;
    }
    return bad ? PLAAC_E_INVALID : PLAAC_OK;
} catch (...) {
    return PLAAC_E_NOMEM;
}

int plaac_unpack_host(const uint32_t* words, int64_t first, int64_t count, uint8_t* codes)
{
    if (first < 0 || count < 0 || (count > 0 && (!words || !codes))) return PLAAC_E_INVALID;
    int64_t r = first;
    const int64_t end = first + count;
    while (r < end) {
        const int64_t w = r / kPer;
        uint32_t v = words[w];
        int k = (int)(r - w * kPer);
        for (int i = 0; i < k; i++) v /= 22u;
        for (; k < kPer && r < end; k++, r++) {
            *codes++ = (uint8_t)(v % 22u);
            v /= 22u;
        }
    }
    return PLAAC_OK;
}

}  // extern "C"
