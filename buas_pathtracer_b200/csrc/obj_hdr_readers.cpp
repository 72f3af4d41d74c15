// Wavefront OBJ and Radiance HDR (RGBE, new-style RLE) readers -- SURVEY 8f rank 3: the two data formats on the input
// side of the path.  Behaviour follows the reference's parse_obj (Raytracer/assets.cpp:187-400) and parse_hdr
// (:411-600) input for input, quirks included, so that a file gives the same triangle list / texel array on both sides
// (tests/test_assets.py compares with the reference's own parsers):
//   OBJ  * any line starting with 'v' adds a vertex; only "vn" / "vt" are told apart ("vp ..." becomes a (0,0,0) vertex);
//        * components come from strtof, face indices from strtol with base 0 ("010" is octal, "0x10" hex), both of which
//          skip white space INCLUDING newlines, so a short "v 1 2" line continues into the next one;
//        * index 0 is a null vertex, negative indices are relative to the current count, faces of up to 32 corners are
//          fanned around their first corner, clockwise winding swaps corners 0 and 2;
//        * texture coordinates are three floats per corner like everything else.
//   HDR  * header lines until the first empty line; only -Y/+Y <h> -X/+X <w>; every scanline must carry the 02 02 marker;
//          the scanline-length bytes are read as SIGNED chars, so widths whose low byte is >= 128 are rejected;
//        * rows are stored bottom-up for "-Y" (the first scanline goes to the LAST row); exponents <= 9 decode to 0;
//          value = 2^(e-136) * (mantissa + 0.5).
// Where the reference would read out of bounds or spin forever on malformed input, this code returns an error instead.
#include "host_scene.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

using bpt::set_error;

struct bpt_obj {
    uint32_t triangle_count = 0;
    bool has_normals = false, has_texcoords = false;
    std::vector<float> positions, normals, texcoords;      // 9 floats per triangle
};

namespace {

struct Float3 { float e[3]; };

bool read_file(const char* path, std::string* out) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), f)) > 0) out->append(buf, got);
    fclose(f);
    return true;
}

bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n'; }

} // namespace

extern "C" {

bpt_obj* bpt_parse_obj(const char* text, int32_t winding) {
    if (!text) { set_error("bpt_parse_obj: null text"); return nullptr; }
    if (winding != 0 && winding != 1) { set_error("bpt_parse_obj: winding must be 0 (clockwise) or 1 (counter-clockwise)"); return nullptr; }
    std::vector<Float3> pools[3];                          // vertices, texture coordinates, normals; [0] is the null entry
    for (auto& p : pools) p.push_back(Float3{{0, 0, 0}});
    std::vector<float> tris[3];                            // fanned triangles per pool

    const char* at = text;
    while (*at) {
        while (*at && is_blank(*at)) ++at;
        const char* line_end = at;
        while (*line_end && *line_end != '\r' && *line_end != '\n') ++line_end;
        const char* next_line = line_end;
        if (*next_line == '\r') ++next_line;
        if (*next_line == '\n') ++next_line;

        char command = *at;
        if (command) ++at;
        if (command == 'v') {
            int pool = 0;
            if (*at == 'n') { ++at; pool = 2; }
            else if (*at == 't') { ++at; pool = 1; }
            Float3 v = {{0, 0, 0}};
            for (int i = 0; i < 3; ++i) {
                char* end;
                float e = strtof(at, &end);
                if (end != at) v.e[i] = e;
                at = end;
            }
            pools[pool].push_back(v);
        } else if (command == 'f') {
            uint32_t idx[3][32];
            uint32_t cnt[3] = {0, 0, 0};
            for (;;) {
                const char* corner_start = at;
                for (int k = 0; k < 3; ++k) {
                    if (cnt[k] >= 32) { set_error("bpt_parse_obj: too many vertices for face"); return nullptr; }
                    char* end;
                    long long index = strtol(at, &end, 0);
                    if (index < 0) {
                        index = (long long)pools[k].size() + index;
                        if (index <= 0) { set_error("bpt_parse_obj: relative index reaches before the first element"); return nullptr; }
                    }
                    if (end != at) {
                        if (index >= (long long)pools[k].size()) { set_error("bpt_parse_obj: index %lld out of range", index); return nullptr; }
                        idx[k][cnt[k]++] = (uint32_t)index;
                    }
                    at = end;
                    if (*at == '/') ++at;
                    else { while (*at == ' ') ++at; break; }
                }
                if (at == line_end) break;
                if (at == corner_start) { set_error("bpt_parse_obj: malformed face"); return nullptr; }   // the reference would spin here
                if (at > line_end) break;            // strtol skipped a newline: the reference's `at == line_end` test would never hit again
            }
            int a = 0, b = 1, c = 2;
            if (winding == 0) { a = 2; c = 0; }
            for (int k = 0; k < 3; ++k) {
                if (!cnt[k]) continue;
                if (cnt[k] < 3) { set_error("bpt_parse_obj: not enough vertices to make a face"); return nullptr; }
                for (uint32_t i = 1; i + 1 < cnt[k]; ++i) {
                    const Float3* corner[3];
                    corner[a] = &pools[k][idx[k][0]];
                    corner[b] = &pools[k][idx[k][i]];
                    corner[c] = &pools[k][idx[k][i + 1]];
                    for (int q = 0; q < 3; ++q) tris[k].insert(tris[k].end(), corner[q]->e, corner[q]->e + 3);
                }
            }
        }
        at = next_line;                              // unconditionally, like the reference (even if strtof looked further)
    }

    size_t n = tris[0].size()/9;
    if (!tris[1].empty() && tris[1].size()/9 != n) { set_error("bpt_parse_obj: texture coordinates don't match triangles"); return nullptr; }
    if (!tris[2].empty() && tris[2].size()/9 != n) { set_error("bpt_parse_obj: normals don't match triangles"); return nullptr; }
    bpt_obj* o = new bpt_obj();
    o->triangle_count = (uint32_t)n;
    // the reference sets has_* from the POOL sizes (assets.cpp:361-367) and then copies from the fanned arrays; when a file
    // lists normals that no face uses that copy reads nothing valid, so here the flag also requires fanned data
    o->has_texcoords = pools[1].size() > 1 && !tris[1].empty();
    o->has_normals = pools[2].size() > 1 && !tris[2].empty();
    o->positions.swap(tris[0]);
    if (o->has_texcoords) o->texcoords.swap(tris[1]);
    if (o->has_normals) o->normals.swap(tris[2]);
    return o;
}

bpt_obj* bpt_load_obj(const char* path, int32_t winding) {
    std::string text;
    if (!path || !read_file(path, &text)) { set_error("bpt_load_obj: cannot read '%s'", path ? path : "(null)"); return nullptr; }
    return bpt_parse_obj(text.c_str(), winding);
}

void bpt_obj_free(bpt_obj* o) { delete o; }
uint32_t bpt_obj_triangle_count(const bpt_obj* o) { return o ? o->triangle_count : 0; }
const float* bpt_obj_positions(const bpt_obj* o) { return o && o->triangle_count ? o->positions.data() : nullptr; }
const float* bpt_obj_normals(const bpt_obj* o) { return o && o->has_normals ? o->normals.data() : nullptr; }
const float* bpt_obj_texcoords(const bpt_obj* o) { return o && o->has_texcoords ? o->texcoords.data() : nullptr; }

uint32_t bpt_create_mesh_from_obj(bpt_scene* s, const bpt_obj* o) {
    if (!s || !o || o->triangle_count == 0) { set_error("bpt_create_mesh_from_obj: empty mesh"); return 0xFFFFFFFFu; }
    // load_mesh (raytracer.cpp:148-158) builds file meshes with the midpoint-split BVH, not the binned-SAH one
    return bpt_create_mesh_ex(s, o->triangle_count, o->positions.data(), o->has_normals ? o->normals.data() : nullptr, BPT_BVH_MIDPOINT_SPLIT);
}

int bpt_parse_hdr(const char* data, size_t size, uint32_t* out_w, uint32_t* out_h, float* pixels) {
    if (!data || !out_w || !out_h) { set_error("bpt_parse_hdr: null argument"); return BPT_ERR_ARG; }
    const char* at = data;
    const char* end = data + size;
    auto left = [&]() { return (size_t)(end - at); };
    auto match = [&](const char* word) {
        const char* p = at;
        while (p < end && *p == ' ') ++p;
        size_t len = strlen(word);
        if ((size_t)(end - p) >= len && memcmp(p, word, len) == 0) { at = p + len; return true; }
        return false;
    };
    // header: lines until an empty one (only FORMAT / PRIMARIES are even looked at by the reference; neither changes the result)
    bool header_done = false;
    while (at < end && *at) {
        if (*at == '\n') { ++at; header_done = true; break; }
        if (match("FORMAT")) {
            if (!match("=")) { set_error("bpt_parse_hdr: malformed header"); return BPT_ERR_ARG; }
        } else if (match("PRIMARIES")) {
            if (!match("=")) { set_error("bpt_parse_hdr: malformed header"); return BPT_ERR_ARG; }
        }
        while (at < end && *at && *at != '\n') ++at;
        if (at < end && *at == '\n') ++at;
    }
    if (!header_done || at >= end || !*at) { set_error("bpt_parse_hdr: unexpected end of file while parsing header"); return BPT_ERR_ARG; }

    auto parse_u32 = [&](uint32_t* v) {
        std::string tmp(at, (size_t)std::min<size_t>(left(), 32));
        char* e;
        unsigned long x = strtoul(tmp.c_str(), &e, 0);
        if (e == tmp.c_str()) return false;
        at += e - tmp.c_str();
        *v = (uint32_t)x;
        return true;
    };
    int x_advance, y_advance;
    uint32_t w = 0, h = 0;
    if (match("+Y")) y_advance = 1; else if (match("-Y")) y_advance = -1;
    else { set_error("bpt_parse_hdr: failed to parse resolution string (+/-Y)"); return BPT_ERR_ARG; }
    if (!parse_u32(&h)) { set_error("bpt_parse_hdr: failed to parse vertical resolution"); return BPT_ERR_ARG; }
    if (match("+X")) x_advance = 1; else if (match("-X")) x_advance = -1;
    else { set_error("bpt_parse_hdr: failed to parse resolution string (+/-X)"); return BPT_ERR_ARG; }
    if (!parse_u32(&w)) { set_error("bpt_parse_hdr: failed to parse horizontal resolution"); return BPT_ERR_ARG; }
    if (at >= end || *at++ != '\n') { set_error("bpt_parse_hdr: expected newline after resolution string"); return BPT_ERR_ARG; }
    if (!w || !h) { set_error("bpt_parse_hdr: malformed resolution"); return BPT_ERR_ARG; }
    *out_w = w; *out_h = h;
    if (!pixels) return BPT_OK;                    // size query

    std::vector<uint8_t> rgbe((size_t)w*h*4, 0);
    ptrdiff_t row = 0;
    if (x_advance < 0) row += (ptrdiff_t)w - 1;
    if (y_advance < 0) row += (ptrdiff_t)w*((ptrdiff_t)h - 1);
    for (uint32_t y = 0; y < h; ++y) {
        if (left() < 4) { set_error("bpt_parse_hdr: truncated scanline"); return BPT_ERR_ARG; }
        // (at[0] << 8) | at[1] on (signed) chars, truncated to 16 bits -- see the note at the top
        uint16_t signature = (uint16_t)(((int)(signed char)at[0] << 8) | (int)(signed char)at[1]);
        at += 2;
        if (signature != 0x0202) { set_error("bpt_parse_hdr: .hdr format unsupported"); return BPT_ERR_UNSUPPORTED; }
        uint16_t scanline_length = (uint16_t)(((int)(signed char)at[0] << 8) | (int)(signed char)at[1]);
        at += 2;
        if (scanline_length != w) { set_error("bpt_parse_hdr: scanline length did not match image width"); return BPT_ERR_ARG; }
        for (int channel = 0; channel < 4; ++channel) {
            ptrdiff_t dst = row;
            for (uint32_t x = 0; x < w;) {
                if (left() < 1) { set_error("bpt_parse_hdr: truncated scanline"); return BPT_ERR_ARG; }
                uint8_t code = (uint8_t)*at++;
                uint32_t count = code > 128 ? (uint32_t)(code & 127) : (uint32_t)code;
                if (count == 0 || x + count > w) { set_error("bpt_parse_hdr: corrupt run-length data"); return BPT_ERR_ARG; }
                if (code > 128) {
                    if (left() < 1) { set_error("bpt_parse_hdr: truncated scanline"); return BPT_ERR_ARG; }
                    uint8_t value = (uint8_t)*at++;
                    for (uint32_t i = 0; i < count; ++i, ++x, dst += x_advance) rgbe[(size_t)dst*4 + channel] = value;
                } else {
                    if (left() < count) { set_error("bpt_parse_hdr: truncated scanline"); return BPT_ERR_ARG; }
                    for (uint32_t i = 0; i < count; ++i, ++x, dst += x_advance) rgbe[(size_t)dst*4 + channel] = (uint8_t)*at++;
                }
            }
        }
        row += (ptrdiff_t)y_advance*(ptrdiff_t)w;
    }
    for (size_t i = 0; i < (size_t)w*h; ++i) {                                             // decode_radiance_color :411-421
        const uint8_t* c = &rgbe[i*4];
        float* o = pixels + i*3;
        o[0] = o[1] = o[2] = 0.0f;
        if (c[3] > 9) {
            uint32_t bits = (uint32_t)(c[3] - 9) << 23;
            float mul;
            memcpy(&mul, &bits, 4);
            o[0] = mul*((float)c[0] + 0.5f);
            o[1] = mul*((float)c[1] + 0.5f);
            o[2] = mul*((float)c[2] + 0.5f);
        }
    }
    return BPT_OK;
}

int bpt_load_skydome_hdr(bpt_scene* s, const char* path) {
    std::string data;
    if (!s || !path || !read_file(path, &data)) { set_error("bpt_load_skydome_hdr: cannot read '%s'", path ? path : "(null)"); return BPT_ERR_ARG; }
    uint32_t w = 0, h = 0;
    int rc = bpt_parse_hdr(data.data(), data.size(), &w, &h, nullptr);
    if (rc) return rc;
    std::vector<float> pixels((size_t)w*h*3);
    rc = bpt_parse_hdr(data.data(), data.size(), &w, &h, pixels.data());
    if (rc) return rc;
    return bpt_set_skydome(s, w, h, pixels.data());
}

} // extern "C"
