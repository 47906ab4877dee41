// Baseline JPEG ingest for the headless host (the reference loads its .jpg textures through
// stb_image, src/graphics/Image.cpp:10-41 — an un-pinned third party whose integer IDCT and chroma filter differ from
// libjpeg's by +-1..2 LSB, so parity is unpinned either way). This decoder follows the Independent JPEG Group's
// published pipeline instead — the slow-but-accurate integer IDCT (13-bit constants, two passes), "fancy" triangle
// chroma upsampling (h2v1 / h2v2 / h1v2) and the 16-bit fixed-point YCbCr -> RGB conversion — which is what PIL
// (libjpeg-turbo, bit-exact with the IJG integer path) produces, so tests/test_cpp_host.py can compare byte for byte.
// Supported: 8-bit baseline / extended sequential (SOF0 / SOF1) and progressive (SOF2: spectral selection + successive
// approximation, ITU-T T.81 annex G, the coefficient rules of IJG jdphuff.c) Huffman JPEG, 1 or 3 components, luma
// sampling 1x1, 2x1, 1x2, 2x2 with 1x1 chroma, interleaved or per-component scans, restart intervals, Adobe transform
// flag 0 (RGB). Arithmetic-coded, lossless, hierarchical, 12-bit and CMYK files are refused loudly.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "texture.h"

namespace rbhost {

namespace {

[[noreturn]] void jfail(const std::string& name, const std::string& why) {
    throw std::runtime_error("Could not load image at path: " + name + ": " + why);
}

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff {
    bool present = false;
    int maxcode[18];     // largest code of each length (-1 if none), [17] = sentinel
    int valptr[17];
    int mincode[17];
    uint8_t vals[256];
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int width = 0, height = 0;            // downsampled size in samples
    int stride = 0, rows = 0;             // plane size, padded to whole MCUs
    int64_t pred = 0;
    std::vector<uint8_t> plane;
    int blocksX = 0, blocksY = 0;         // progressive only: quantised coefficients of every block, zigzag order
    std::vector<int> coef;
};

struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint32_t acc = 0;
    int bits = 0;
    bool hitMarker = false;
    void fill() {
        while (bits <= 24) {
            int byte = 0;
            if (!hitMarker && p < end) {
                byte = *p;
                if (byte == 0xFF) {
                    if (p + 1 < end && p[1] == 0x00) p += 2;
                    else { hitMarker = true; byte = 0; }      // a marker: feed zeros, do not consume it
                } else p++;
            }
            acc |= uint32_t(byte) << (24 - bits);
            bits += 8;
        }
    }
    int get(int n) {
        if (n == 0) return 0;
        if (bits < n) fill();
        const int v = int(acc >> (32 - n));
        acc <<= n;
        bits -= n;
        return v;
    }
    int peek16() { if (bits < 16) fill(); return int(acc >> 16); }
    void skip(int n) { acc <<= n; bits -= n; }
    void reset() { acc = 0; bits = 0; hitMarker = false; }
};

int decode_symbol(BitReader& br, const Huff& h, const std::string& name) {
    const int look = br.peek16();
    for (int len = 1; len <= 16; len++) {
        const int code = look >> (16 - len);
        if (h.maxcode[len] >= 0 && code <= h.maxcode[len] && code >= h.mincode[len]) {
            br.skip(len);
            return h.vals[h.valptr[len] + code - h.mincode[len]];
        }
    }
    jfail(name, "corrupt JPEG data: bad Huffman code");
}

inline int extend(int v, int n) { return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v; }

inline uint8_t clamp8(int v) { return uint8_t(v < 0 ? 0 : v > 255 ? 255 : v); }

// quantised coefficients are 16-bit in a valid file (JCOEF); corrupt data saturates instead of wrapping
inline int clamp16(int64_t v) { return int(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }

// de-quantised coefficients of a valid 8-bit file stay within +-2^15; corrupt data may not
inline int clamp_coef(int64_t v) { return int(v < -(1 << 24) ? -(1 << 24) : v > (1 << 24) ? (1 << 24) : v); }

// jpeg_idct_islow (IJG jidctint.c): CONST_BITS 13, PASS1_BITS 2. 64-bit intermediates: identical to the 32-bit original
// on every valid file, and free of signed overflow on corrupt ones (coefficients are clamped to +-2^24 by the caller).
void idct_islow(const int* coef, uint8_t* out, int stride) {
    using I = int64_t;
    auto descale = [](I x, int n) -> I { return (x + (I(1) << (n - 1))) >> n; };
    auto clamp8 = [](I v) -> uint8_t { return uint8_t(v < 0 ? 0 : v > 255 ? 255 : v); };
    enum { C_0_298 = 2446, C_0_390 = 3196, C_0_541 = 4433, C_0_765 = 6270, C_0_899 = 7373, C_1_175 = 9633, C_1_501 = 12299,
           C_1_847 = 15137, C_1_961 = 16069, C_2_053 = 16819, C_2_562 = 20995, C_3_072 = 25172 };
    I ws[64];
    for (int c = 0; c < 8; c++) {
        const int* in = coef + c;
        I z2 = in[16], z3 = in[48];
        I z1 = (z2 + z3) * C_0_541;
        I tmp2 = z1 + z3 * (-C_1_847);
        I tmp3 = z1 + z2 * C_0_765;
        z2 = in[0]; z3 = in[32];
        I tmp0 = (z2 + z3) * 8192;
        I tmp1 = (z2 - z3) * 8192;
        const I tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = in[56]; tmp1 = in[40]; tmp2 = in[24]; tmp3 = in[8];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
        I z4 = tmp1 + tmp3;
        const I z5 = (z3 + z4) * C_1_175;
        tmp0 *= C_0_298; tmp1 *= C_2_053; tmp2 *= C_3_072; tmp3 *= C_1_501;
        z1 *= -C_0_899; z2 *= -C_2_562; z3 *= -C_1_961; z4 *= -C_0_390;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        I* w = ws + c;
        w[0] = descale(tmp10 + tmp3, 11); w[56] = descale(tmp10 - tmp3, 11);
        w[8] = descale(tmp11 + tmp2, 11); w[48] = descale(tmp11 - tmp2, 11);
        w[16] = descale(tmp12 + tmp1, 11); w[40] = descale(tmp12 - tmp1, 11);
        w[24] = descale(tmp13 + tmp0, 11); w[32] = descale(tmp13 - tmp0, 11);
    }
    for (int r = 0; r < 8; r++) {
        const I* w = ws + 8 * r;
        I z2 = w[2], z3 = w[6];
        I z1 = (z2 + z3) * C_0_541;
        I tmp2 = z1 + z3 * (-C_1_847);
        I tmp3 = z1 + z2 * C_0_765;
        I tmp0 = (w[0] + w[4]) * 8192;
        I tmp1 = (w[0] - w[4]) * 8192;
        const I tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
        I z4 = tmp1 + tmp3;
        const I z5 = (z3 + z4) * C_1_175;
        tmp0 *= C_0_298; tmp1 *= C_2_053; tmp2 *= C_3_072; tmp3 *= C_1_501;
        z1 *= -C_0_899; z2 *= -C_2_562; z3 *= -C_1_961; z4 *= -C_0_390;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        uint8_t* o = out + r * stride;
        o[0] = clamp8(descale(tmp10 + tmp3, 18) + 128); o[7] = clamp8(descale(tmp10 - tmp3, 18) + 128);
        o[1] = clamp8(descale(tmp11 + tmp2, 18) + 128); o[6] = clamp8(descale(tmp11 - tmp2, 18) + 128);
        o[2] = clamp8(descale(tmp12 + tmp1, 18) + 128); o[5] = clamp8(descale(tmp12 - tmp1, 18) + 128);
        o[3] = clamp8(descale(tmp13 + tmp0, 18) + 128); o[4] = clamp8(descale(tmp13 - tmp0, 18) + 128);
    }
}

// jdsample.c, h2v1_fancy_upsample: one input row of n samples -> 2n samples
void upsample_h2_row(const uint8_t* in, int n, uint8_t* out) {
    if (n == 1) { out[0] = out[1] = in[0]; return; }
    int v = in[0];
    out[0] = uint8_t(v);
    out[1] = uint8_t((v * 3 + in[1] + 2) >> 2);
    for (int i = 1; i < n - 1; i++) {
        v = in[i] * 3;
        out[2 * i] = uint8_t((v + in[i - 1] + 1) >> 2);
        out[2 * i + 1] = uint8_t((v + in[i + 1] + 2) >> 2);
    }
    v = in[n - 1];
    out[2 * (n - 1)] = uint8_t((v * 3 + in[n - 2] + 1) >> 2);
    out[2 * (n - 1) + 1] = uint8_t(v);
}

// h2v2_fancy_upsample: near = the input row this output row belongs to, far = its neighbour above / below
void upsample_h2v2_row(const uint8_t* near, const uint8_t* far, int n, uint8_t* out) {
    if (n == 1) { out[0] = out[1] = uint8_t((near[0] * 3 + far[0] + 2) >> 2); return; }     // (4 * colsum + 8) >> 4
    int thiscol = near[0] * 3 + far[0], nextcol = near[1] * 3 + far[1], lastcol;
    out[0] = uint8_t((thiscol * 4 + 8) >> 4);
    out[1] = uint8_t((thiscol * 3 + nextcol + 7) >> 4);
    lastcol = thiscol; thiscol = nextcol;
    for (int i = 1; i < n - 1; i++) {
        nextcol = near[i + 1] * 3 + far[i + 1];
        out[2 * i] = uint8_t((thiscol * 3 + lastcol + 8) >> 4);
        out[2 * i + 1] = uint8_t((thiscol * 3 + nextcol + 7) >> 4);
        lastcol = thiscol; thiscol = nextcol;
    }
    out[2 * (n - 1)] = uint8_t((thiscol * 3 + lastcol + 8) >> 4);
    out[2 * (n - 1) + 1] = uint8_t((thiscol * 4 + 7) >> 4);
}

}  // namespace

bool looks_like_jpeg(const uint8_t* data, size_t size) { return size >= 3 && data[0] == 0xFF && data[1] == 0xD8 && data[2] == 0xFF; }

Image8 decode_jpeg_rgba8(const uint8_t* data, size_t size, bool flip, const std::string& name) {
    if (!looks_like_jpeg(data, size)) jfail(name, "not a JPEG file");
    uint16_t quant[4][64];
    bool haveQuant[4] = {false, false, false, false};
    Huff dc[4], ac[4];
    std::vector<Component> comps;
    int width = 0, height = 0, hmax = 1, vmax = 1, restartInterval = 0;
    bool sawFrame = false, adobe = false, decodedAny = false, progressive = false;
    int adobeTransform = -1;
    size_t pos = 2;
    auto be16 = [&](size_t at) -> int { if (at + 2 > size) jfail(name, "truncated JPEG"); return (data[at] << 8) | data[at + 1]; };

    for (;;) {
        // next marker
        while (pos < size && data[pos] != 0xFF) pos++;
        while (pos < size && data[pos] == 0xFF) pos++;
        if (pos >= size) break;
        const int marker = data[pos++];
        if (marker == 0xD9) break;                              // EOI
        if (marker == 0x00 || marker == 0x01 || (marker >= 0xD0 && marker <= 0xD7)) continue;   // stuffed byte, TEM, RSTn
        const int len = be16(pos);
        if (len < 2 || pos + size_t(len) > size) jfail(name, "truncated JPEG segment");
        const uint8_t* seg = data + pos + 2;
        const int n = len - 2;
        if (marker == 0xDB) {                                    // DQT
            int i = 0;
            while (i < n) {
                const int pq = seg[i] >> 4, tq = seg[i] & 15;
                i++;
                if (tq > 3) jfail(name, "bad quantisation table id");
                if (i + (pq ? 128 : 64) > n) jfail(name, "truncated DQT");
                for (int k = 0; k < 64; k++) {
                    quant[tq][kZigzag[k]] = pq ? uint16_t((seg[i] << 8) | seg[i + 1]) : seg[i];
                    i += pq ? 2 : 1;
                }
                haveQuant[tq] = true;
            }
        } else if (marker == 0xC4) {                             // DHT
            int i = 0;
            while (i + 17 <= n) {
                const int tc = seg[i] >> 4, th = seg[i] & 15;
                if (tc > 1 || th > 3) jfail(name, "bad Huffman table id");
                Huff& h = tc ? ac[th] : dc[th];
                int counts[17], total = 0;
                for (int l = 1; l <= 16; l++) { counts[l] = seg[i + l]; total += counts[l]; }
                i += 17;
                if (total > 256 || i + total > n) jfail(name, "truncated DHT");
                std::memcpy(h.vals, seg + i, size_t(total));
                i += total;
                int code = 0, k = 0;
                for (int l = 1; l <= 16; l++) {
                    h.valptr[l] = k;
                    h.mincode[l] = code;
                    if (counts[l]) { code += counts[l]; h.maxcode[l] = code - 1; k += counts[l]; }
                    else h.maxcode[l] = -1;
                    code <<= 1;
                }
                h.present = true;
            }
        } else if (marker == 0xC0 || marker == 0xC1 || marker == 0xC2) {           // SOF0 / SOF1 / SOF2
            if (sawFrame) jfail(name, "more than one frame");
            progressive = marker == 0xC2;
            if (n < 6) jfail(name, "truncated SOF");
            if (seg[0] != 8) jfail(name, "only 8-bit JPEG is supported");
            height = (seg[1] << 8) | seg[2];
            width = (seg[3] << 8) | seg[4];
            const int nc = seg[5];
            if (width == 0 || height == 0 || width > 65535 || height > 65535) jfail(name, "unsupported dimensions");
            if (int64_t(width) * int64_t(height) > (int64_t(1) << 28)) jfail(name, "image too large");      // 268 Mpixel
            if (nc != 1 && nc != 3) jfail(name, "only grayscale and 3-component JPEG are supported");
            if (n < 6 + 3 * nc) jfail(name, "truncated SOF");
            comps.resize(size_t(nc));
            for (int c = 0; c < nc; c++) {
                Component& k = comps[size_t(c)];
                k.id = seg[6 + 3 * c];
                k.h = seg[7 + 3 * c] >> 4; k.v = seg[7 + 3 * c] & 15;
                k.tq = seg[8 + 3 * c];
                if (k.h < 1 || k.h > 2 || k.v < 1 || k.v > 2 || k.tq > 3) jfail(name, "unsupported sampling factors");
                hmax = std::max(hmax, k.h); vmax = std::max(vmax, k.v);
            }
            if (nc == 1) { comps[0].h = comps[0].v = 1; hmax = vmax = 1; }
            if (nc == 3 && (comps[1].h != 1 || comps[1].v != 1 || comps[2].h != 1 || comps[2].v != 1 || comps[0].h != hmax || comps[0].v != vmax))
                jfail(name, "unsupported sampling factors");
            const int mcusX = (width + 8 * hmax - 1) / (8 * hmax), mcusY = (height + 8 * vmax - 1) / (8 * vmax);
            for (Component& k : comps) {
                k.width = (width * k.h + hmax - 1) / hmax;
                k.height = (height * k.v + vmax - 1) / vmax;
                k.stride = mcusX * k.h * 8;
                k.rows = mcusY * k.v * 8;
                k.plane.assign(size_t(k.stride) * size_t(k.rows), 128);
                if (progressive) {
                    k.blocksX = k.stride / 8; k.blocksY = k.rows / 8;
                    k.coef.assign(size_t(k.blocksX) * size_t(k.blocksY) * 64, 0);
                }
            }
            sawFrame = true;
        } else if ((marker >= 0xC3 && marker <= 0xCF) && marker != 0xC4 && marker != 0xC8 && marker != 0xCC)
            jfail(name, "unsupported JPEG coding process");
        else if (marker == 0xCC) jfail(name, "arithmetic-coded JPEG is not supported");
        else if (marker == 0xDD) { if (n < 2) jfail(name, "truncated DRI"); restartInterval = (seg[0] << 8) | seg[1]; }
        else if (marker == 0xEE && n >= 12 && std::memcmp(seg, "Adobe", 5) == 0) { adobe = true; adobeTransform = seg[11]; }
        else if (marker == 0xDA) {                               // SOS + entropy-coded data
            if (!sawFrame) jfail(name, "scan before frame header");
            if (n < 1) jfail(name, "truncated SOS");
            const int ns = seg[0];
            if (ns < 1 || ns > int(comps.size()) || n < 1 + 2 * ns + 3) jfail(name, "bad SOS");
            // progression parameters (T.81 G.1.1.1): band Ss..Se, previous / current point transform Ah / Al
            const int Ss = seg[1 + 2 * ns], Se = seg[2 + 2 * ns], Ah = seg[3 + 2 * ns] >> 4, Al = seg[3 + 2 * ns] & 15;
            if (progressive && (Ss > Se || Se > 63 || Al > 13 || (Ss == 0 && Se != 0) || (Ss > 0 && ns != 1) || (Ah != 0 && Al != Ah - 1)))
                jfail(name, "bad progression parameters");
            const bool needDc = !progressive || (Ss == 0 && Ah == 0), needAc = !progressive || Ss > 0;
            std::vector<Component*> scan;
            for (int s = 0; s < ns; s++) {
                const int cid = seg[1 + 2 * s];
                Component* found = nullptr;
                for (Component& k : comps) if (k.id == cid) found = &k;
                if (!found) jfail(name, "scan refers to an unknown component");
                found->td = seg[2 + 2 * s] >> 4; found->ta = seg[2 + 2 * s] & 15;
                if (found->td > 3 || found->ta > 3 || (needDc && !dc[found->td].present) || (needAc && !ac[found->ta].present))
                    jfail(name, "missing Huffman table");
                if (!progressive && !haveQuant[found->tq]) jfail(name, "missing quantisation table");
                for (Component* seen : scan) if (seen == found) jfail(name, "component listed twice in a scan");
                scan.push_back(found);
            }
            const uint8_t* sp = seg + n;
            BitReader br{sp, data + size};
            for (Component& k : comps) k.pred = 0;
            const bool interleaved = ns > 1;
            int unitsX, unitsY;
            if (interleaved) { unitsX = (width + 8 * hmax - 1) / (8 * hmax); unitsY = (height + 8 * vmax - 1) / (8 * vmax); }
            else { unitsX = (scan[0]->width + 7) / 8; unitsY = (scan[0]->height + 7) / 8; }
            int untilRestart = restartInterval, expectRst = 0;
            int block[64];
            int eobrun = 0;                                     // progressive AC scans: blocks left in the end-of-band run
            // one block of a progressive scan (jdphuff.c decode_mcu_DC_first / DC_refine / AC_first / AC_refine)
            auto progressive_block = [&](Component* k, int bx, int by) {
                int scratch[64] = {0};
                int* b = (bx < k->blocksX && by < k->blocksY) ? &k->coef[(size_t(by) * size_t(k->blocksX) + size_t(bx)) * 64] : scratch;
                const int p1 = 1 << Al, m1 = -(1 << Al);
                if (Ss == 0) {
                    if (Ah == 0) {
                        const int t = decode_symbol(br, dc[k->td], name);
                        if (t > 11) jfail(name, "corrupt JPEG data: bad DC size");
                        k->pred += t ? extend(br.get(t), t) : 0;
                        b[0] = clamp16(k->pred * p1);
                    } else if (br.get(1)) b[0] |= p1;
                    return;
                }
                const Huff& h = ac[k->ta];
                if (Ah == 0) {                                  // first pass over the band
                    if (eobrun > 0) { eobrun--; return; }
                    for (int i = Ss; i <= Se; i++) {
                        const int rs = decode_symbol(br, h, name);
                        const int r = rs >> 4, s = rs & 15;
                        if (s) {
                            i += r;
                            if (i > 63) jfail(name, "corrupt JPEG data: coefficient index out of range");
                            b[i] = clamp16(int64_t(extend(br.get(s), s)) * p1);
                        } else if (r == 15) i += 15;
                        else {                                  // EOBr: this block and eobrun more end here
                            eobrun = (1 << r) - 1;
                            if (r) eobrun += br.get(r);
                            break;
                        }
                    }
                    return;
                }
                // refinement: one more bit for the coefficients that are already non-zero, new +-1 coefficients between them
                auto correct = [&](int& c) {
                    if (br.get(1) && (c & p1) == 0) c += c >= 0 ? p1 : m1;
                };
                int i = Ss;
                if (eobrun == 0) {
                    for (; i <= Se; i++) {
                        const int rs = decode_symbol(br, h, name);
                        int r = rs >> 4, s = rs & 15;
                        if (s) s = br.get(1) ? p1 : m1;         // the size must be 1: a newly non-zero coefficient
                        else if (r != 15) {
                            eobrun = 1 << r;
                            if (r) eobrun += br.get(r);
                            break;                              // the rest of the band is handled below
                        }
                        // skip r still-zero coefficients, refining the non-zero ones on the way
                        for (; i <= Se; i++) {
                            if (b[i] != 0) correct(b[i]);
                            else if (--r < 0) break;
                        }
                        if (s) {
                            if (i > 63) jfail(name, "corrupt JPEG data: coefficient index out of range");
                            b[i] = s;
                        }
                    }
                }
                if (eobrun > 0) {
                    for (; i <= Se; i++) if (b[i] != 0) correct(b[i]);
                    eobrun--;
                }
            };
            for (int uy = 0; uy < unitsY; uy++)
                for (int ux = 0; ux < unitsX; ux++) {
                    if (restartInterval && untilRestart == 0) {
                        // byte-align, expect RSTn
                        const uint8_t* q = br.p;
                        while (q < data + size && !(q[0] == 0xFF && q + 1 < data + size && q[1] >= 0xD0 && q[1] <= 0xD7)) q++;
                        if (q >= data + size) jfail(name, "corrupt JPEG data: missing restart marker");
                        if (q[1] != 0xD0 + expectRst) jfail(name, "corrupt JPEG data: restart markers out of order");
                        expectRst = (expectRst + 1) & 7;
                        br.p = q + 2;
                        br.reset();
                        for (Component& k : comps) k.pred = 0;
                        eobrun = 0;
                        untilRestart = restartInterval;
                    }
                    for (Component* k : scan) {
                        const int bw = interleaved ? k->h : 1, bh = interleaved ? k->v : 1;
                        for (int by = 0; by < bh; by++)
                            for (int bx = 0; bx < bw; bx++) {
                                if (progressive) { progressive_block(k, ux * bw + bx, uy * bh + by); continue; }
                                std::memset(block, 0, sizeof block);
                                const uint16_t* q = quant[k->tq];
                                int t = decode_symbol(br, dc[k->td], name);
                                if (t > 11) jfail(name, "corrupt JPEG data: bad DC size");
                                const int diff = t ? extend(br.get(t), t) : 0;
                                k->pred += diff;
                                block[0] = clamp_coef(k->pred * int64_t(q[0]));
                                for (int i = 1; i < 64;) {
                                    const int rs = decode_symbol(br, ac[k->ta], name);
                                    const int r = rs >> 4, s = rs & 15;
                                    if (s == 0) {
                                        if (r != 15) break;      // end of block
                                        i += 16;
                                        continue;
                                    }
                                    i += r;
                                    if (i > 63) jfail(name, "corrupt JPEG data: coefficient index out of range");
                                    const int zz = kZigzag[i];
                                    block[zz] = clamp_coef(int64_t(extend(br.get(s), s)) * int64_t(q[zz]));
                                    i++;
                                }
                                const int px = (ux * bw + bx) * 8, py = (uy * bh + by) * 8;
                                if (px + 8 <= k->stride && py + 8 <= k->rows)
                                    idct_islow(block, &k->plane[size_t(py) * size_t(k->stride) + size_t(px)], k->stride);
                            }
                    }
                    if (restartInterval) untilRestart--;
                }
            decodedAny = true;
            pos = size_t(br.p - data);
            continue;
        } else if (marker == 0xD8) jfail(name, "unexpected SOI");
        pos += size_t(len);
    }
    if (!sawFrame || !decodedAny) jfail(name, "no image data");
    if (progressive) {                                          // all scans are in: de-quantise and transform every block
        int block[64];
        for (Component& k : comps) {
            if (!haveQuant[k.tq]) jfail(name, "missing quantisation table");
            const uint16_t* q = quant[k.tq];
            for (int by = 0; by < k.blocksY; by++)
                for (int bx = 0; bx < k.blocksX; bx++) {
                    const int* b = &k.coef[(size_t(by) * size_t(k.blocksX) + size_t(bx)) * 64];
                    for (int i = 0; i < 64; i++) block[kZigzag[i]] = clamp_coef(int64_t(b[i]) * int64_t(q[kZigzag[i]]));
                    idct_islow(block, &k.plane[size_t(by) * 8 * size_t(k.stride) + size_t(bx) * 8], k.stride);
                }
        }
    }

    Image8 img;
    img.width = width; img.height = height;
    img.rgba.assign(size_t(width) * size_t(height) * 4, 255);
    const bool gray = comps.size() == 1;
    const bool ycc = !gray && !(adobe && adobeTransform == 0);
    // chroma rows at full resolution, one output row at a time
    std::vector<uint8_t> up[3];
    for (auto& u : up) u.assign(size_t(width) + 2 * 8 * 2 + 4, 0);
    for (int y = 0; y < height; y++) {
        const uint8_t* rows[3] = {nullptr, nullptr, nullptr};
        for (size_t c = 0; c < comps.size(); c++) {
            const Component& k = comps[c];
            const int sx = hmax / k.h, sy = vmax / k.v;          // 1 or 2
            if (sx == 1 && sy == 1) { rows[c] = &k.plane[size_t(y) * size_t(k.stride)]; continue; }
            const int n = k.width;
            auto row = [&](int r) { r = r < 0 ? 0 : r >= k.height ? k.height - 1 : r; return &k.plane[size_t(r) * size_t(k.stride)]; };
            if (sy == 1) {                                      // h2v1
                if (n > 2) upsample_h2_row(row(y), n, up[c].data());
                else for (int x = 0; x < n; x++) up[c][size_t(2 * x)] = up[c][size_t(2 * x + 1)] = row(y)[x];
            } else {
                const int r = y >> 1;
                const uint8_t* near = row(r);
                const uint8_t* far = (y & 1) ? row(r + 1) : row(r - 1);
                if (sx == 2) {                                  // h2v2
                    if (n > 2) upsample_h2v2_row(near, far, n, up[c].data());
                    else for (int x = 0; x < n; x++) up[c][size_t(2 * x)] = up[c][size_t(2 * x + 1)] = near[x];
                } else {                                        // h1v2: (3 * near + far + bias) >> 2, bias 1 above / 2 below
                    const int bias = (y & 1) ? 2 : 1;
                    for (int x = 0; x < n; x++) up[c][size_t(x)] = uint8_t((near[x] * 3 + far[x] + bias) >> 2);
                }
            }
            rows[c] = up[c].data();
        }
        uint8_t* o = &img.rgba[size_t(flip ? height - 1 - y : y) * size_t(width) * 4];
        for (int x = 0; x < width; x++, o += 4) {
            if (gray) { o[0] = o[1] = o[2] = rows[0][x]; continue; }
            const int Y = rows[0][x], cb = rows[1][x] - 128, cr = rows[2][x] - 128;
            if (!ycc) { o[0] = rows[0][x]; o[1] = rows[1][x]; o[2] = rows[2][x]; continue; }
            // jdcolor.c: SCALEBITS 16, FIX(1.40200) = 91881, FIX(1.77200) = 116130, FIX(0.71414) = 46802, FIX(0.34414) = 22554
            o[0] = clamp8(Y + ((91881 * cr + 32768) >> 16));
            o[1] = clamp8(Y + ((-22554 * cb + 32768 - 46802 * cr) >> 16));
            o[2] = clamp8(Y + ((116130 * cb + 32768) >> 16));
        }
    }
    return img;
}

}  // namespace rbhost
