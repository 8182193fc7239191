// png.cpp -- see png.hpp.  Written from the PNG (ISO 15948) and DEFLATE (RFC 1951) specifications.
#include "png.hpp"

#include <cstdlib>
#include <cstring>

namespace luzhost {

namespace {

struct BitReader {
    const uint8_t* p;
    size_t n, pos = 0;
    uint32_t acc = 0;
    int cnt = 0;
    bool overrun = false;
    uint32_t bits(int k) { // LSB first
        while (cnt < k) {
            uint32_t b = 0;
            if (pos < n)
                b = p[pos++];
            else
                overrun = true;
            acc |= b << cnt;
            cnt += 8;
        }
        const uint32_t v = acc & ((k == 32) ? 0xFFFFFFFFu : ((1u << k) - 1u));
        acc >>= k;
        cnt -= k;
        return v;
    }
    void align() {
        acc = 0;
        cnt = 0;
    }
};

// canonical Huffman decoding table: counts per length + symbols sorted by code
struct Huffman {
    uint16_t count[16];
    uint16_t symbol[288];
    bool build(const uint8_t* lengths, int n) {
        memset(count, 0, sizeof count);
        for (int i = 0; i < n; i++) count[lengths[i]]++;
        count[0] = 0;
        int left = 1;
        for (int len = 1; len < 16; len++) {
            left <<= 1;
            left -= count[len];
            if (left < 0) return false; // over-subscribed
        }
        uint16_t offs[16];
        offs[1] = 0;
        for (int len = 1; len < 15; len++) offs[len + 1] = offs[len] + count[len];
        for (int i = 0; i < n; i++)
            if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
        return true;
    }
    int decode(BitReader& br) const {
        int code = 0, first = 0, index = 0;
        for (int len = 1; len < 16; len++) {
            code |= (int)br.bits(1);
            const int c = count[len];
            if (code - c < first) return symbol[index + (code - first)];
            index += c;
            first += c;
            first <<= 1;
            code <<= 1;
            if (br.overrun) return -1;
        }
        return -1;
    }
};

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint16_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint16_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

bool inflate_block(BitReader& br, const Huffman& lit, const Huffman& dist, std::vector<uint8_t>& out) {
    while (true) {
        const int sym = lit.decode(br);
        if (sym < 0) return false;
        if (sym < 256) {
            out.push_back((uint8_t)sym);
        } else if (sym == 256) {
            return true;
        } else {
            const int li = sym - 257;
            if (li >= 29) return false;
            const int len = kLenBase[li] + (int)br.bits(kLenExtra[li]);
            const int ds = dist.decode(br);
            if (ds < 0 || ds >= 30) return false;
            const size_t d = kDistBase[ds] + br.bits(kDistExtra[ds]);
            if (d > out.size()) return false;
            const size_t start = out.size() - d;
            for (int k = 0; k < len; k++) out.push_back(out[start + k]);
        }
        if (br.overrun) return false;
    }
}

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    if (pb <= pc) return b;
    return c;
}

} // namespace

bool inflate_zlib(const uint8_t* data, size_t size, std::vector<uint8_t>& out, std::string& err) {
    if (size < 6 || (data[0] & 0x0F) != 8 || ((data[0] << 8) | data[1]) % 31 != 0 || (data[1] & 0x20)) {
        err = "not a zlib stream";
        return false;
    }
    BitReader br{data + 2, size - 2};
    bool last = false;
    while (!last) {
        last = br.bits(1) != 0;
        const uint32_t type = br.bits(2);
        if (type == 0) {
            br.align();
            if (br.pos + 4 > br.n) { err = "truncated stored block"; return false; }
            const uint32_t len = br.p[br.pos] | (br.p[br.pos + 1] << 8), nlen = br.p[br.pos + 2] | (br.p[br.pos + 3] << 8);
            br.pos += 4;
            if ((len ^ 0xFFFFu) != nlen || br.pos + len > br.n) { err = "bad stored block"; return false; }
            out.insert(out.end(), br.p + br.pos, br.p + br.pos + len);
            br.pos += len;
        } else if (type == 1 || type == 2) {
            Huffman lit, dist;
            uint8_t lengths[320];
            if (type == 1) {
                int i = 0;
                for (; i < 144; i++) lengths[i] = 8;
                for (; i < 256; i++) lengths[i] = 9;
                for (; i < 280; i++) lengths[i] = 7;
                for (; i < 288; i++) lengths[i] = 8;
                lit.build(lengths, 288);
                for (i = 0; i < 30; i++) lengths[i] = 5;
                dist.build(lengths, 30);
            } else {
                const int hlit = (int)br.bits(5) + 257, hdist = (int)br.bits(5) + 1, hclen = (int)br.bits(4) + 4;
                static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                uint8_t cl[19] = {0};
                for (int i = 0; i < hclen; i++) cl[order[i]] = (uint8_t)br.bits(3);
                Huffman clh;
                if (hlit > 286 || hdist > 30 || !clh.build(cl, 19)) { err = "bad code length code"; return false; }
                int i = 0;
                while (i < hlit + hdist) {
                    const int sym = clh.decode(br);
                    if (sym < 0) { err = "bad code lengths"; return false; }
                    if (sym < 16) {
                        lengths[i++] = (uint8_t)sym;
                    } else {
                        int rep, val = 0;
                        if (sym == 16) {
                            if (i == 0) { err = "repeat without previous length"; return false; }
                            val = lengths[i - 1];
                            rep = 3 + (int)br.bits(2);
                        } else if (sym == 17) {
                            rep = 3 + (int)br.bits(3);
                        } else {
                            rep = 11 + (int)br.bits(7);
                        }
                        if (i + rep > hlit + hdist) { err = "too many code lengths"; return false; }
                        while (rep--) lengths[i++] = (uint8_t)val;
                    }
                }
                if (!lit.build(lengths, hlit) || !dist.build(lengths + hlit, hdist)) { err = "bad Huffman code"; return false; }
            }
            if (!inflate_block(br, lit, dist, out)) { err = "corrupt deflate data"; return false; }
        } else {
            err = "reserved block type";
            return false;
        }
        if (br.overrun) { err = "truncated deflate data"; return false; }
    }
    return true;
}

bool decode_png(const uint8_t* data, size_t size, std::vector<uint8_t>& rgba, int& width, int& height, std::string& err,
                std::vector<uint16_t>* deep) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (size < 8 || memcmp(data, sig, 8) != 0) { err = "not a PNG"; return false; }
    size_t off = 8;
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat;
    uint8_t pal[256][4];
    int pal_n = 0;
    bool has_key = false;
    uint16_t key[3] = {0, 0, 0};
    for (int i = 0; i < 256; i++) pal[i][0] = pal[i][1] = pal[i][2] = 0, pal[i][3] = 255;
    bool end = false;
    while (!end && off + 12 <= size) {
        const uint32_t len = be32(data + off);
        const uint8_t* type = data + off + 4;
        const uint8_t* body = data + off + 8;
        if (off + 12 + (size_t)len > size) { err = "truncated chunk"; return false; }
        if (!memcmp(type, "IHDR", 4)) {
            if (len < 13) { err = "bad IHDR"; return false; }
            w = be32(body);
            h = be32(body + 4);
            depth = body[8];
            ctype = body[9];
            interlace = body[12];
        } else if (!memcmp(type, "PLTE", 4)) {
            pal_n = (int)(len / 3);
            for (int i = 0; i < pal_n && i < 256; i++) pal[i][0] = body[3 * i], pal[i][1] = body[3 * i + 1], pal[i][2] = body[3 * i + 2];
        } else if (!memcmp(type, "tRNS", 4)) {
            if (ctype == 3) {
                for (uint32_t i = 0; i < len && i < 256; i++) pal[i][3] = body[i];
            } else if (ctype == 0 && len >= 2) {
                has_key = true;
                key[0] = (uint16_t)((body[0] << 8) | body[1]);
            } else if (ctype == 2 && len >= 6) {
                has_key = true;
                for (int k = 0; k < 3; k++) key[k] = (uint16_t)((body[2 * k] << 8) | body[2 * k + 1]);
            }
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!memcmp(type, "IEND", 4)) {
            end = true;
        }
        off += 12 + (size_t)len;
    }
    if (!w || !h || ctype < 0) { err = "missing IHDR"; return false; }
    if (interlace > 1) { err = "bad interlace method"; return false; }
    int channels;
    switch (ctype) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: err = "bad colour type"; return false;
    }
    if ((ctype == 2 || ctype == 4 || ctype == 6) && depth != 8 && depth != 16) { err = "bad bit depth"; return false; }
    if (ctype == 3 && depth == 16) { err = "bad bit depth"; return false; }
    if (depth != 1 && depth != 2 && depth != 4 && depth != 8 && depth != 16) { err = "bad bit depth"; return false; }
    if ((uint64_t)w * h > (1ull << 28)) { err = "image too large"; return false; }
    std::vector<uint8_t> raw;
    raw.reserve(((size_t)w * channels * depth / 8 + 2) * h);
    if (!inflate_zlib(idat.data(), idat.size(), raw, err)) return false;
    const size_t bpp = (size_t)((channels * depth + 7) / 8);
    rgba.assign((size_t)w * h * 4, 255);
    if (deep) deep->clear();
    if (deep && depth == 16) deep->assign((size_t)w * h * 4, 0xFFFF);
    static const int scale_table[9] = {0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01};

    // one (sub)image of pw x ph pixels starting at raw[pos]: unfilter, then scatter its pixels to (x0 + x*dx, y0 + y*dy)
    auto decode_pass = [&](size_t& pos, uint32_t pw, uint32_t ph, uint32_t x0, uint32_t y0, uint32_t dx, uint32_t dy) -> bool {
        if (!pw || !ph) return true;
        const size_t stride = ((size_t)pw * channels * depth + 7) / 8;
        if (raw.size() < pos + (stride + 1) * ph) { err = "not enough image data"; return false; }
        std::vector<uint8_t> prev(stride, 0), cur(stride);
        for (uint32_t y = 0; y < ph; y++) {
            const uint8_t* row = raw.data() + pos + (stride + 1) * y;
            const int filter = row[0];
            memcpy(cur.data(), row + 1, stride);
            for (size_t i = 0; i < stride; i++) {
                const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
                int v = cur[i];
                switch (filter) {
                    case 0: break;
                    case 1: v += a; break;
                    case 2: v += b; break;
                    case 3: v += (a + b) >> 1; break;
                    case 4: v += paeth(a, b, c); break;
                    default: err = "bad filter type"; return false;
                }
                cur[i] = (uint8_t)v;
            }
            // expand to RGBA8 the way stb_image does when asked for 4 channels of 8 bits: 16-bit samples keep their
            // high byte, the tRNS colour key is compared at full depth, sub-byte grey is scaled to 0..255
            const uint8_t* r = cur.data();
            for (uint32_t x = 0; x < pw; x++) {
                uint8_t* o = rgba.data() + ((size_t)(y0 + y * dy) * w + (x0 + x * dx)) * 4;
                auto sample = [&](size_t k) -> uint32_t { // k-th sample of the row at full depth (8 or 16 bit)
                    return depth == 16 ? (uint32_t)((r[2 * k] << 8) | r[2 * k + 1]) : r[k];
                };
                auto to8 = [&](uint32_t v) -> uint8_t { return (uint8_t)(depth == 16 ? (v >> 8) : v); };
                if (ctype == 0 || ctype == 3) {
                    uint32_t v;
                    if (depth >= 8) {
                        v = sample(x);
                    } else {
                        const int per = 8 / depth, shift = (per - 1 - (int)(x % per)) * depth;
                        v = (uint32_t)(r[x / per] >> shift) & ((1u << depth) - 1u);
                    }
                    if (ctype == 3) {
                        o[0] = pal[v & 255][0], o[1] = pal[v & 255][1], o[2] = pal[v & 255][2], o[3] = pal[v & 255][3];
                    } else {
                        const bool transparent = has_key && v == key[0];
                        o[0] = o[1] = o[2] = depth < 8 ? (uint8_t)(v * scale_table[depth]) : to8(v);
                        o[3] = transparent ? 0 : 255;
                    }
                } else if (ctype == 2) {
                    const uint32_t R = sample(3 * x), G = sample(3 * x + 1), B = sample(3 * x + 2);
                    o[0] = to8(R), o[1] = to8(G), o[2] = to8(B);
                    o[3] = (has_key && R == key[0] && G == key[1] && B == key[2]) ? 0 : 255;
                } else if (ctype == 4) {
                    o[0] = o[1] = o[2] = to8(sample(2 * x));
                    o[3] = to8(sample(2 * x + 1));
                } else {
                    for (int k = 0; k < 4; k++) o[k] = to8(sample(4 * x + k));
                }
                if (deep && depth == 16) { // the same pixel at full depth (stb's 16-bit API)
                    uint16_t* q = deep->data() + ((size_t)(y0 + y * dy) * w + (x0 + x * dx)) * 4;
                    if (ctype == 0) {
                        q[0] = q[1] = q[2] = (uint16_t)sample(x);
                        q[3] = (has_key && sample(x) == key[0]) ? 0 : 0xFFFF;
                    } else if (ctype == 2) {
                        for (int k = 0; k < 3; k++) q[k] = (uint16_t)sample(3 * x + k);
                        q[3] = o[3] ? 0xFFFF : 0;
                    } else if (ctype == 4) {
                        q[0] = q[1] = q[2] = (uint16_t)sample(2 * x);
                        q[3] = (uint16_t)sample(2 * x + 1);
                    } else {
                        for (int k = 0; k < 4; k++) q[k] = (uint16_t)sample(4 * x + k);
                    }
                }
            }
            prev.swap(cur);
        }
        pos += (stride + 1) * ph;
        return true;
    };
    size_t pos = 0;
    if (!interlace) {
        if (!decode_pass(pos, w, h, 0, 0, 1, 1)) return false;
    } else { // Adam7: seven reduced images, each filtered on its own
        static const uint32_t xs[7] = {0, 4, 0, 2, 0, 1, 0}, ys[7] = {0, 0, 4, 0, 2, 0, 1};
        static const uint32_t dxs[7] = {8, 8, 4, 4, 2, 2, 1}, dys[7] = {8, 8, 8, 4, 4, 2, 2};
        for (int p = 0; p < 7; p++) {
            const uint32_t pw = (w > xs[p]) ? (w - xs[p] + dxs[p] - 1) / dxs[p] : 0;
            const uint32_t ph = (h > ys[p]) ? (h - ys[p] + dys[p] - 1) / dys[p] : 0;
            if (!decode_pass(pos, pw, ph, xs[p], ys[p], dxs[p], dys[p])) return false;
        }
    }
    width = (int)w;
    height = (int)h;
    return true;
}

} // namespace luzhost
