/*
 * jpeg.cpp — baseline / extended-sequential JPEG decode for glTF textures.
 *
 * The reference loads images through tinygltf -> stb_image (deps/sf_libs, third-party, version 2.x)
 * with 4 requested channels (src/scene/scene.cpp:374-388).  Texel values feed shading directly
 * (rt.rgen:97-130), so "identical results" needs the same bytes: this file restates stb_image's
 * published arithmetic — the jidctint-derived 12-bit integer IDCT with 2 guard bits between the passes,
 * its triangle-filter chroma upsampling (3:1 / 9:3:3:1 weights), its 20-bit fixed-point YCbCr->RGB and
 * its rule for when three components are RGB rather than YCbCr — in independent code.  Pinned against
 * the reference's own decoder by tests/golden (SHA-256 of every decoded texture).
 *
 * Supported: 8-bit precision, 1 or 3 components, Huffman baseline / extended sequential (SOF0/SOF1) and
 * progressive (SOF2, T.81 Annex G: spectral selection + successive approximation), interleaved and
 * per-component scans, restart intervals, any sampling factors 1..4 (2x1, 1x2, 2x2 filtered; other
 * ratios replicate, like stb).  Arithmetic coding, lossless and 4-component files are reported as
 * unsupported.
 */
#include <cstring>

#include "scene.h"

namespace gpurt {
namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff {
    /* canonical code table: for each length 1..16 the first code, the index of its first symbol, count */
    int32_t first_code[17], first_sym[17], count[17];
    uint8_t sym[256];
    bool present = false;
    bool build(const uint8_t counts[16], const uint8_t* symbols, int n_symbols) {
        int code = 0, k = 0;
        for(int len = 1; len <= 16; len++) {
            count[len] = counts[len - 1];
            first_code[len] = code;
            first_sym[len] = k;
            code += count[len];
            k += count[len];
            if(code > (1 << len)) return false;
            code <<= 1;
        }
        if(k != n_symbols || k > 256) return false;
        std::memcpy(sym, symbols, (size_t)k);
        present = true;
        return true;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int x = 0, y = 0;   /* size in samples */
    int w2 = 0, h2 = 0; /* allocated size (whole MCUs) */
    int pred = 0;
    std::vector<uint8_t> data;
    std::vector<short> coef; /* progressive: 64 coefficients per block, (w2/8) blocks per row */
};

struct Reader {
    const uint8_t* p;
    size_t n, pos = 0;
    uint32_t bits = 0;
    int nbits = 0;
    int marker = 0; /* marker met inside entropy-coded data; zero bits are fed after it */
    void reset_bits() { bits = 0, nbits = 0, marker = 0; }
    void fill() {
        while(nbits <= 24) {
            int b = 0;
            if(!marker && pos < n) {
                b = p[pos++];
                if(b == 0xFF) {
                    int c = pos < n ? p[pos++] : 0xD9;
                    while(c == 0xFF) c = pos < n ? p[pos++] : 0xD9; /* fill bytes */
                    if(c != 0) {
                        marker = c;
                        b = 0;
                    }
                }
            }
            bits |= (uint32_t)b << (24 - nbits);
            nbits += 8;
        }
    }
    int get(int k) { /* k <= 16 */
        if(k == 0) return 0;
        if(nbits < k) fill();
        int v = (int)(bits >> (32 - k));
        bits <<= k;
        nbits -= k;
        return v;
    }
    int decode(const Huff& h) {
        if(nbits < 16) fill();
        int code = 0;
        for(int len = 1; len <= 16; len++) {
            code = (code << 1) | (int)((bits >> (32 - len)) & 1u);
            int off = code - h.first_code[len];
            if(off >= 0 && off < h.count[len]) {
                bits <<= len;
                nbits -= len;
                return h.sym[h.first_sym[len] + off];
            }
        }
        return -1;
    }
    /* T.81 F.2.2.1 EXTEND */
    int receive_extend(int s) {
        int v = get(s);
        return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
    }
};

inline uint8_t clamp8(int x) { return (uint8_t)(x < 0 ? 0 : x > 255 ? 255 : x); }

/* One 8-point pass of the "islow" integer IDCT (Loeffler-Ligtenberg-Moschytz) with 12-bit constants:
 * even part in x[0..3], odd part in t[0..3]; outputs are x[k] +- t[3-k]. */
inline int fix(float c) { return (int)(c * 4096 + 0.5); } /* float constant, scaled exactly, rounded in double */
struct Pass {
    int x0, x1, x2, x3, t0, t1, t2, t3;
};
/* 32-bit arithmetic that wraps: a valid stream never overflows here and decodes as before; a corrupt one (coefficients far
 * outside the 11-bit range) gets wrapped garbage pixels instead of undefined behaviour */
inline int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
inline int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
inline int wmul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
inline Pass idct8(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7) {
    static const int c0541 = fix(0.5411961f), c1847 = fix(-1.847759065f), c0765 = fix(0.765366865f),
                     c1175 = fix(1.175875602f), c0298 = fix(0.298631336f), c2053 = fix(2.053119869f),
                     c3072 = fix(3.072711026f), c1501 = fix(1.501321110f), c0899 = fix(-0.899976223f),
                     c2562 = fix(-2.562915447f), c1961 = fix(-1.961570560f), c0390 = fix(-0.390180644f);
    Pass r;
    int p1 = wmul(wadd(s2, s6), c0541);
    int e2 = wadd(p1, wmul(s6, c1847)), e3 = wadd(p1, wmul(s2, c0765));
    int e0 = wmul(wadd(s0, s4), 4096), e1 = wmul(wsub(s0, s4), 4096);
    r.x0 = wadd(e0, e3), r.x3 = wsub(e0, e3), r.x1 = wadd(e1, e2), r.x2 = wsub(e1, e2);
    int o0 = s7, o1 = s5, o2 = s3, o3 = s1;
    int q3 = wadd(o0, o2), q4 = wadd(o1, o3), q1 = wadd(o0, o3), q2 = wadd(o1, o2);
    int q5 = wmul(wadd(q3, q4), c1175);
    o0 = wmul(o0, c0298), o1 = wmul(o1, c2053), o2 = wmul(o2, c3072), o3 = wmul(o3, c1501);
    q1 = wadd(q5, wmul(q1, c0899)), q2 = wadd(q5, wmul(q2, c2562));
    q3 = wmul(q3, c1961), q4 = wmul(q4, c0390);
    r.t3 = wadd(wadd(o3, q1), q4), r.t2 = wadd(wadd(o2, q2), q3), r.t1 = wadd(wadd(o1, q2), q4), r.t0 = wadd(wadd(o0, q1), q3);
    return r;
}

void idct_block(uint8_t* out, int stride, const short d[64]) {
    int v[64];
    for(int c = 0; c < 8; c++) { /* columns: keep 2 extra bits */
        const short* s = d + c;
        if(!s[8] && !s[16] && !s[24] && !s[32] && !s[40] && !s[48] && !s[56]) {
            int dc = s[0] * 4;
            for(int r = 0; r < 8; r++) v[r * 8 + c] = dc;
            continue;
        }
        Pass p = idct8(s[0], s[8], s[16], s[24], s[32], s[40], s[48], s[56]);
        p.x0 = wadd(p.x0, 512), p.x1 = wadd(p.x1, 512), p.x2 = wadd(p.x2, 512), p.x3 = wadd(p.x3, 512);
        v[0 * 8 + c] = wadd(p.x0, p.t3) >> 10, v[7 * 8 + c] = wsub(p.x0, p.t3) >> 10;
        v[1 * 8 + c] = wadd(p.x1, p.t2) >> 10, v[6 * 8 + c] = wsub(p.x1, p.t2) >> 10;
        v[2 * 8 + c] = wadd(p.x2, p.t1) >> 10, v[5 * 8 + c] = wsub(p.x2, p.t1) >> 10;
        v[3 * 8 + c] = wadd(p.x3, p.t0) >> 10, v[4 * 8 + c] = wsub(p.x3, p.t0) >> 10;
    }
    for(int r = 0; r < 8; r++) { /* rows: remove 2^17, round, level shift by 128 */
        const int* s = v + r * 8;
        uint8_t* o = out + (size_t)r * stride;
        Pass p = idct8(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7]);
        const int bias = 65536 + (128 << 17);
        p.x0 = wadd(p.x0, bias), p.x1 = wadd(p.x1, bias), p.x2 = wadd(p.x2, bias), p.x3 = wadd(p.x3, bias);
        o[0] = clamp8(wadd(p.x0, p.t3) >> 17), o[7] = clamp8(wsub(p.x0, p.t3) >> 17);
        o[1] = clamp8(wadd(p.x1, p.t2) >> 17), o[6] = clamp8(wsub(p.x1, p.t2) >> 17);
        o[2] = clamp8(wadd(p.x2, p.t1) >> 17), o[5] = clamp8(wsub(p.x2, p.t1) >> 17);
        o[3] = clamp8(wadd(p.x3, p.t0) >> 17), o[4] = clamp8(wsub(p.x3, p.t0) >> 17);
    }
}

/* chroma upsampling of one output row; `near` is the closer source row, `far` the other one */
const uint8_t* upsample_row(uint8_t* out, const uint8_t* near, const uint8_t* far, int w, int hs, int vs) {
    if(hs == 1 && vs == 1) return near;
    if(hs == 1 && vs == 2) {
        for(int i = 0; i < w; i++) out[i] = (uint8_t)((3 * near[i] + far[i] + 2) >> 2);
        return out;
    }
    if(hs == 2 && vs == 1) {
        if(w == 1) {
            out[0] = out[1] = near[0];
            return out;
        }
        out[0] = near[0];
        out[1] = (uint8_t)((near[0] * 3 + near[1] + 2) >> 2);
        int i;
        for(i = 1; i < w - 1; i++) {
            int n = 3 * near[i] + 2;
            out[2 * i] = (uint8_t)((n + near[i - 1]) >> 2);
            out[2 * i + 1] = (uint8_t)((n + near[i + 1]) >> 2);
        }
        out[2 * i] = (uint8_t)((near[w - 2] * 3 + near[w - 1] + 2) >> 2);
        out[2 * i + 1] = near[w - 1];
        return out;
    }
    if(hs == 2 && vs == 2) {
        if(w == 1) {
            out[0] = out[1] = (uint8_t)((3 * near[0] + far[0] + 2) >> 2);
            return out;
        }
        int t1 = 3 * near[0] + far[0];
        out[0] = (uint8_t)((t1 + 2) >> 2);
        for(int i = 1; i < w; i++) {
            int t0 = t1;
            t1 = 3 * near[i] + far[i];
            out[2 * i - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
            out[2 * i] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
        }
        out[2 * w - 1] = (uint8_t)((t1 + 2) >> 2);
        return out;
    }
    for(int i = 0; i < w; i++) /* other ratios: replicate horizontally, nearest row vertically */
        for(int j = 0; j < hs; j++) out[i * hs + j] = near[i];
    return out;
}

inline int ycc_fix(float c) { return ((int)(c * 4096.0f + 0.5f)) << 8; }

} // namespace

bool decode_jpeg(const std::vector<uint8_t>& f, Texture& out, std::string& err) {
    Reader rd{f.data(), f.size()};
    auto fail = [&](const char* m) {
        err = std::string("JPEG: ") + m;
        return false;
    };
    if(f.size() < 4 || f[0] != 0xFF || f[1] != 0xD8) return fail("no SOI");
    rd.pos = 2;
    uint16_t quant[4][64] = {};
    Huff hdc[4], hac[4];
    Component comp[3];
    int n_comp = 0, img_w = 0, img_h = 0, hmax = 1, vmax = 1, mcu_x = 0, mcu_y = 0;
    int restart = 0;
    bool jfif = false, frame = false, saw_scan = false, progressive = false;
    int adobe_transform = -1, rgb_ids = 0;

    for(;;) {
        /* next marker */
        int m = rd.marker;
        rd.marker = 0;
        if(!m) {
            if(rd.pos + 1 >= rd.n) break;
            if(rd.p[rd.pos++] != 0xFF) continue;
            m = rd.p[rd.pos++];
            while(m == 0xFF && rd.pos < rd.n) m = rd.p[rd.pos++];
            if(m == 0) continue;
        }
        if(m == 0xD9) break; /* EOI */
        if(m >= 0xD0 && m <= 0xD7) continue;
        if(rd.pos + 2 > rd.n) return fail("truncated");
        size_t len = ((size_t)rd.p[rd.pos] << 8) | rd.p[rd.pos + 1];
        if(len < 2 || rd.pos + len > rd.n) return fail("bad segment length");
        const uint8_t* s = rd.p + rd.pos + 2;
        size_t sl = len - 2;
        rd.pos += len;
        if(m == 0xE0) {
            if(sl >= 5 && !std::memcmp(s, "JFIF\0", 5)) jfif = true;
        } else if(m == 0xEE) {
            if(sl >= 12 && !std::memcmp(s, "Adobe\0", 6)) adobe_transform = s[11];
        } else if(m == 0xDB) {
            while(sl > 0) {
                int pq = s[0] >> 4, tq = s[0] & 15;
                size_t need = 1 + (pq ? 128 : 64);
                if(pq > 1 || tq > 3 || sl < need) return fail("bad DQT");
                for(int i = 0; i < 64; i++)
                    quant[tq][kZigzag[i]] = pq ? (uint16_t)((s[1 + 2 * i] << 8) | s[2 + 2 * i]) : s[1 + i];
                s += need, sl -= need;
            }
        } else if(m == 0xC4) {
            while(sl > 0) {
                if(sl < 17) return fail("bad DHT");
                int tc = s[0] >> 4, th = s[0] & 15, total = 0;
                for(int i = 0; i < 16; i++) total += s[1 + i];
                if(tc > 1 || th > 3 || sl < (size_t)(17 + total)) return fail("bad DHT");
                if(!(tc ? hac : hdc)[th].build(s + 1, s + 17, total)) return fail("bad Huffman table");
                s += 17 + total, sl -= 17 + total;
            }
        } else if(m == 0xDD) {
            if(sl != 2) return fail("bad DRI");
            restart = (s[0] << 8) | s[1];
        } else if(m == 0xC0 || m == 0xC1 || m == 0xC2) {
            progressive = m == 0xC2;
            if(frame) return fail("two frames");
            if(sl < 6 || s[0] != 8) return fail("only 8-bit precision is supported");
            img_h = (s[1] << 8) | s[2], img_w = (s[3] << 8) | s[4], n_comp = s[5];
            if(!img_w || !img_h) return fail("empty image");
            if((uint64_t)img_w * img_h > (1ull << 28)) return fail("image larger than 2^28 pixels");
            /* untrusted dimensions: a coded 8 x 8 block takes at least one bit, so a file of n bytes holds at most 8 n blocks =
             * 512 n pixels — a header that promises more is refused before planes of that size are allocated */
            if((uint64_t)img_w * img_h > 512ull * f.size()) return fail("image dimensions exceed what the file can encode");
            if(n_comp != 1 && n_comp != 3) return fail("only 1- and 3-component files are supported");
            if(sl < (size_t)(6 + 3 * n_comp)) return fail("bad SOF");
            static const char rgb[3] = {'R', 'G', 'B'};
            for(int i = 0; i < n_comp; i++) {
                Component& c = comp[i];
                c.id = s[6 + 3 * i], c.h = s[7 + 3 * i] >> 4, c.v = s[7 + 3 * i] & 15, c.tq = s[8 + 3 * i];
                if(n_comp == 3 && c.id == rgb[i]) rgb_ids++;
                if(c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return fail("bad sampling factors");
                hmax = std::max(hmax, c.h), vmax = std::max(vmax, c.v);
            }
            for(int i = 0; i < n_comp; i++)
                if(hmax % comp[i].h || vmax % comp[i].v) return fail("non-integer sampling ratio");
            mcu_x = (img_w + 8 * hmax - 1) / (8 * hmax), mcu_y = (img_h + 8 * vmax - 1) / (8 * vmax);
            for(int i = 0; i < n_comp; i++) {
                Component& c = comp[i];
                c.x = (img_w * c.h + hmax - 1) / hmax, c.y = (img_h * c.v + vmax - 1) / vmax;
                c.w2 = mcu_x * c.h * 8, c.h2 = mcu_y * c.v * 8;
                c.data.assign((size_t)c.w2 * c.h2, 0);
                if(progressive) c.coef.assign((size_t)c.w2 * c.h2, 0);
            }
            frame = true;
        } else if(m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
            return fail("lossless / hierarchical / arithmetic-coded files are not supported");
        } else if(m == 0xDA) {
            if(!frame) return fail("scan before frame");
            int ns = sl ? s[0] : 0;
            if(ns < 1 || ns > n_comp || sl < (size_t)(4 + 2 * ns)) return fail("bad SOS");
            int order[3];
            for(int i = 0; i < ns; i++) {
                int which = -1;
                for(int k = 0; k < n_comp; k++)
                    if(comp[k].id == s[1 + 2 * i]) which = k;
                if(which < 0) return fail("bad SOS component");
                comp[which].td = s[2 + 2 * i] >> 4, comp[which].ta = s[2 + 2 * i] & 15;
                if(comp[which].td > 3 || comp[which].ta > 3) return fail("bad SOS table");
                order[i] = which;
            }
            const int ss = s[1 + 2 * ns], se = s[2 + 2 * ns], ah = s[3 + 2 * ns] >> 4, al = s[3 + 2 * ns] & 15;
            if(progressive) {
                if(ss > 63 || se > 63 || ss > se || ah > 13 || al > 13) return fail("bad SOS spectral / approximation fields");
                if(ss != 0 && ns != 1) return fail("interleaved AC scan");
                if(ss == 0 && se != 0) return fail("scan mixes DC and AC");
            }
            int eob_run = 0;
            saw_scan = true;
            rd.reset_bits();
            for(int k = 0; k < n_comp; k++) comp[k].pred = 0;
            int todo = restart ? restart : 0x7fffffff;
            short blk[64];
            auto block = [&](Component& c, int bx, int by) -> bool {
                const Huff &D = hdc[c.td], &A = hac[c.ta];
                if(!D.present || !A.present) return false;
                std::memset(blk, 0, sizeof(blk));
                int t = rd.decode(D);
                if(t < 0 || t > 15) return false;
                c.pred += t ? rd.receive_extend(t) : 0;
                blk[0] = (short)(c.pred * quant[c.tq][0]);
                for(int k = 1; k < 64;) {
                    int rs = rd.decode(A);
                    if(rs < 0) return false;
                    int r = rs >> 4, sz = rs & 15;
                    if(sz == 0) {
                        if(rs != 0xF0) break; /* EOB */
                        k += 16;
                    } else {
                        k += r;
                        if(k > 63) return false;
                        int z = kZigzag[k++];
                        blk[z] = (short)(rd.receive_extend(sz) * quant[c.tq][z]);
                    }
                }
                idct_block(c.data.data() + (size_t)c.w2 * by * 8 + (size_t)bx * 8, c.w2, blk);
                return true;
            };
            /* progressive scans accumulate coefficients (T.81 G.1.2); IDCT happens after the last scan */
            auto prog_block = [&](Component& c, int bx, int by) -> bool {
                short* d = c.coef.data() + 64 * ((size_t)by * (c.w2 >> 3) + bx);
                if(ss == 0) { /* DC: first pass codes the prediction difference, later passes one bit */
                    if(ah == 0) {
                        const Huff& D = hdc[c.td];
                        if(!D.present) return false;
                        std::memset(d, 0, 64 * sizeof(short));
                        int t = rd.decode(D);
                        if(t < 0 || t > 15) return false;
                        c.pred += t ? rd.receive_extend(t) : 0;
                        d[0] = (short)(c.pred * (1 << al));
                    } else if(rd.get(1))
                        d[0] += (short)(1 << al);
                    return true;
                }
                const Huff& A = hac[c.ta];
                if(!A.present) return false;
                if(ah == 0) { /* AC first pass of the band ss..se, with end-of-band runs */
                    if(eob_run) {
                        eob_run--;
                        return true;
                    }
                    for(int k = ss; k <= se;) {
                        int rs = rd.decode(A);
                        if(rs < 0) return false;
                        int r = rs >> 4, sz = rs & 15;
                        if(sz == 0) {
                            if(r < 15) {
                                eob_run = (1 << r) + (r ? rd.get(r) : 0) - 1;
                                break;
                            }
                            k += 16;
                        } else {
                            k += r;
                            if(k > 63) return false;
                            d[kZigzag[k++]] = (short)(rd.receive_extend(sz) * (1 << al));
                        }
                    }
                    return true;
                }
                /* AC refinement: one more bit for known coefficients, new +-1 coefficients in between */
                const short bit = (short)(1 << al);
                auto refine = [&](short& p) {
                    if(rd.get(1) && (p & bit) == 0) p = (short)(p > 0 ? p + bit : p - bit);
                };
                if(eob_run) {
                    eob_run--;
                    for(int k = ss; k <= se; k++) {
                        short& p = d[kZigzag[k]];
                        if(p != 0) refine(p);
                    }
                    return true;
                }
                for(int k = ss; k <= se;) {
                    int rs = rd.decode(A);
                    if(rs < 0) return false;
                    int r = rs >> 4, sz = rs & 15, val = 0;
                    if(sz == 0) {
                        if(r < 15) {
                            eob_run = (1 << r) - 1 + (r ? rd.get(r) : 0);
                            r = 64; /* rest of the band only gets refinement bits */
                        }
                    } else {
                        if(sz != 1) return false;
                        val = rd.get(1) ? bit : -bit;
                    }
                    while(k <= se) {
                        short& p = d[kZigzag[k++]];
                        if(p != 0) refine(p);
                        else {
                            if(r == 0) {
                                p = (short)val;
                                break;
                            }
                            r--;
                        }
                    }
                }
                return true;
            };
            auto after_mcu = [&]() -> bool { /* false: stop this scan */
                if(--todo > 0) return true;
                if(rd.nbits < 24) rd.fill();
                if(!(rd.marker >= 0xD0 && rd.marker <= 0xD7)) return false;
                eob_run = 0;
                rd.reset_bits();
                for(int k = 0; k < n_comp; k++) comp[k].pred = 0;
                todo = restart ? restart : 0x7fffffff;
                return true;
            };
            bool go = true;
            if(ns == 1) { /* non-interleaved: blocks of this component only, no MCU padding */
                Component& c = comp[order[0]];
                int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
                for(int j = 0; j < h && go; j++)
                    for(int i = 0; i < w && go; i++) {
                        if(!(progressive ? prog_block(c, i, j) : block(c, i, j))) return fail("bad entropy-coded data");
                        go = after_mcu();
                    }
            } else {
                for(int j = 0; j < mcu_y && go; j++)
                    for(int i = 0; i < mcu_x && go; i++) {
                        for(int q = 0; q < ns; q++) {
                            Component& c = comp[order[q]];
                            for(int y = 0; y < c.v; y++)
                                for(int x = 0; x < c.h; x++)
                                    if(!(progressive ? prog_block(c, i * c.h + x, j * c.v + y) : block(c, i * c.h + x, j * c.v + y)))
                                        return fail("bad entropy-coded data");
                        }
                        go = after_mcu();
                    }
            }
            /* a marker swallowed by the bit reader is handled by the main loop (rd.marker) */
        }
    }
    if(!frame || !saw_scan) return fail("no image data");
    if(progressive) /* dequantise (16-bit product, as stored) and transform every block that holds image samples */
        for(int k = 0; k < n_comp; k++) {
            Component& c = comp[k];
            for(int by = 0; by < (c.y + 7) >> 3; by++)
                for(int bx = 0; bx < (c.x + 7) >> 3; bx++) {
                    short* d = c.coef.data() + 64 * ((size_t)by * (c.w2 >> 3) + bx);
                    for(int i = 0; i < 64; i++) d[i] = (short)(d[i] * quant[c.tq][i]);
                    idct_block(c.data.data() + (size_t)c.w2 * by * 8 + (size_t)bx * 8, c.w2, d);
                }
        }

    /* three components are RGB when tagged 'R','G','B' or when Adobe says "no transform" without JFIF */
    bool is_rgb = n_comp == 3 && (rgb_ids == 3 || (adobe_transform == 0 && !jfif));
    out.w = (uint32_t)img_w, out.h = (uint32_t)img_h;
    out.rgba.assign((size_t)img_w * img_h * 4, 255);
    struct Up {
        int hs, vs, ystep, w_lores, ypos;
        const uint8_t *line0, *line1;
        std::vector<uint8_t> buf;
    } up[3];
    for(int k = 0; k < n_comp; k++) {
        Up& r = up[k];
        r.hs = hmax / comp[k].h, r.vs = vmax / comp[k].v;
        r.ystep = r.vs >> 1, r.w_lores = (img_w + r.hs - 1) / r.hs, r.ypos = 0;
        r.line0 = r.line1 = comp[k].data.data();
        r.buf.assign((size_t)img_w + 8, 0);
    }
    const int cr_r = ycc_fix(1.40200f), cr_g = ycc_fix(0.71414f), cb_g = ycc_fix(0.34414f), cb_b = ycc_fix(1.77200f);
    for(int j = 0; j < img_h; j++) {
        const uint8_t* row[3];
        for(int k = 0; k < n_comp; k++) {
            Up& r = up[k];
            bool bottom = r.ystep >= (r.vs >> 1);
            row[k] = upsample_row(r.buf.data(), bottom ? r.line1 : r.line0, bottom ? r.line0 : r.line1, r.w_lores, r.hs, r.vs);
            if(++r.ystep >= r.vs) {
                r.ystep = 0;
                r.line0 = r.line1;
                if(++r.ypos < comp[k].y) r.line1 += comp[k].w2;
            }
        }
        uint8_t* o = out.rgba.data() + (size_t)j * img_w * 4;
        if(n_comp == 1) {
            for(int i = 0; i < img_w; i++, o += 4) o[0] = o[1] = o[2] = row[0][i];
        } else if(is_rgb) {
            for(int i = 0; i < img_w; i++, o += 4) o[0] = row[0][i], o[1] = row[1][i], o[2] = row[2][i];
        } else {
            for(int i = 0; i < img_w; i++, o += 4) {
                int yf = (row[0][i] << 20) + (1 << 19);
                int cb = row[1][i] - 128, cr = row[2][i] - 128;
                int r = yf + cr * cr_r;
                int g = yf + cr * -cr_g + (int)((unsigned)(cb * -cb_g) & 0xffff0000u);
                int b = yf + cb * cb_b;
                o[0] = clamp8(r >> 20), o[1] = clamp8(g >> 20), o[2] = clamp8(b >> 20);
            }
        }
    }
    return true;
}

} // namespace gpurt
