// pack2bit.cpp -- host side of the packed 2-bit transport of HP-1 (north_star: "packed 2-bit k-mer windows").
//
// A sequence batch crosses PCIe at 1 byte per base; 99.9 % of those bytes are one of ACGTacgt.  tc_pack2bit turns a
// batch into 2 bits per base (code = (c >> 1) & 3: A/a 0, C/c 1, T/t 2, G/g 3 -- the bit trick of the reference's
// complement table, RabbitSketch/src/MinHash.cpp:1185-1195) plus an exception list (position, byte) for every other
// byte, so that the device can rebuild the exact bytes MinHash::update hashes (it upper-cases them itself,
// MinHash.cpp:472-479: a lower-case base unpacks to its upper-case form, which is what gets hashed anyway).
// AVX2 when the CPU has it (32 bases -> 8 bytes in ~8 instructions), scalar otherwise; the caller runs it on several
// threads over disjoint ranges.
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <immintrin.h>

namespace {

inline bool is_acgt(uint8_t c)
{
    c &= 0xDF;
    return c == 'A' || c == 'C' || c == 'G' || c == 'T';
}

// scalar: bases [i0, i1) (i0 a multiple of 4) -> out[(i - i0) / 4 ...]; exceptions appended (positions relative to `origin`)
size_t pack_scalar(const uint8_t *seq, uint64_t i0, uint64_t i1, uint64_t origin, uint8_t *out, uint32_t *exc_pos, uint8_t *exc_byte,
                   size_t exc_n, size_t exc_cap)
{
    for (uint64_t i = i0; i < i1; i += 4) {
        uint8_t b = 0;
        for (int k = 0; k < 4 && i + k < i1; k++) {
            const uint8_t c = seq[i + k];
            b |= (uint8_t)(((c >> 1) & 3) << (2 * k));
            if (!is_acgt(c)) {
                if (exc_n < exc_cap) { exc_pos[exc_n] = (uint32_t)(i + k - origin); exc_byte[exc_n] = c; }
                exc_n++;
            }
        }
        out[(i - i0) >> 2] = b;
    }
    return exc_n;
}

__attribute__((target("avx2")))
size_t pack_avx2(const uint8_t *seq, uint64_t i0, uint64_t i1, uint64_t origin, uint8_t *out, uint32_t *exc_pos, uint8_t *exc_byte,
                 size_t exc_n, size_t exc_cap)
{
    const __m256i m3 = _mm256_set1_epi8(3), mdf = _mm256_set1_epi8((char)0xDF);
    const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'), cT = _mm256_set1_epi8('T');
    const __m256i w14 = _mm256_set1_epi16(0x0401);           // bytes (1, 4): b0 + 4*b1
    const __m256i w116 = _mm256_set1_epi32(0x00100001);      // words (1, 16): lo + 16*hi
    const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                            0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    uint64_t i = i0;
    for (; i + 32 <= i1; i += 32) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(seq + i));
        const __m256i up = _mm256_and_si256(v, mdf);
        const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(up, cA), _mm256_cmpeq_epi8(up, cC)),
                                           _mm256_or_si256(_mm256_cmpeq_epi8(up, cG), _mm256_cmpeq_epi8(up, cT)));
        const uint32_t okm = (uint32_t)_mm256_movemask_epi8(ok);
        if (okm != 0xFFFFFFFFu) {
            uint32_t bad = ~okm;
            while (bad) {
                const int k = __builtin_ctz(bad);
                if (exc_n < exc_cap) { exc_pos[exc_n] = (uint32_t)(i + k - origin); exc_byte[exc_n] = seq[i + k]; }
                exc_n++;
                bad &= bad - 1;
            }
        }
        const __m256i code = _mm256_and_si256(_mm256_srli_epi16(v, 1), m3);          // per byte: (c >> 1) & 3
        const __m256i p2 = _mm256_maddubs_epi16(code, w14);                           // 16-bit: b0 + 4 b1
        const __m256i p4 = _mm256_madd_epi16(p2, w116);                               // 32-bit: (b0+4b1) + 16 (b2+4b3)
        const __m256i g = _mm256_shuffle_epi8(p4, gather);                            // byte 0 of every dword, per 128-bit lane
        const uint32_t lo = (uint32_t)_mm256_extract_epi32(g, 0), hi = (uint32_t)_mm256_extract_epi32(g, 4);
        const uint64_t w = (uint64_t)lo | ((uint64_t)hi << 32);
        memcpy(out + ((i - i0) >> 2), &w, 8);
    }
    if (i < i1) exc_n = pack_scalar(seq, i, i1, origin, out + ((i - i0) >> 2), exc_pos, exc_byte, exc_n, exc_cap);
    return exc_n;
}

}  // namespace

// Pack bases [i0, i1) of seq (i0 a multiple of 4) into out (2 bits per base, base i at bits 2*(i&3) of byte (i - i0)/4).
// Bytes other than ACGTacgt are reported in exc_pos/exc_byte (positions relative to `origin`), starting at slot 0 of the
// given arrays; returns how many there were (more than exc_cap: the caller's list was too small, nothing past it is stored).
extern "C" size_t tc_pack2bit_range(const uint8_t *seq, uint64_t i0, uint64_t i1, uint64_t origin, uint8_t *out,
                                    uint32_t *exc_pos, uint8_t *exc_byte, size_t exc_cap)
{
    static const bool avx2 = __builtin_cpu_supports("avx2");
    return avx2 ? pack_avx2(seq, i0, i1, origin, out, exc_pos, exc_byte, 0, exc_cap)
                : pack_scalar(seq, i0, i1, origin, out, exc_pos, exc_byte, 0, exc_cap);
}
