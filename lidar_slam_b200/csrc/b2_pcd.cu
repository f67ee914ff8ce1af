// b2_pcd.cu -- PCD v0.7 file I/O for PointXYZI clouds (host code; SURVEY 8(f) row 4: the on-disk format either side
// of the hot path).  The reference reads maps and key frames with pcl::io::loadPCDFile (src/matching/matching.cpp:155,
// src/mapping/loop_closing/loop_closing.cpp:134,286,304) and writes them with pcl::io::savePCDFileBinary
// (src/mapping/back_end/back_end.cpp:194, src/mapping/viewer/viewer.cpp:202,210).  PCL is not available here, so this
// follows the published PCD v0.7 format: an ASCII header (VERSION, FIELDS, SIZE, TYPE, COUNT, WIDTH, HEIGHT, VIEWPOINT,
// POINTS, DATA) followed by ascii rows or packed binary records.  pcl::PointXYZI is stored as the four float32 fields
// x y z intensity = 16 bytes per point, which is exactly the device layout {x,y,z,intensity}: a binary PointXYZI file is
// uploaded without any repacking.  DATA binary_compressed (field-major records in one LZF stream, what
// pcl::io::savePCDFileBinaryCompressed and most third-party map tools write) is read as well; files are written as
// DATA binary only, like the reference.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#include "b2_cloud.cuh"

using namespace b2;

namespace {
struct PcdField { std::string name; int size = 4; char type = 'F'; int count = 1; size_t offset = 0; };

double field_value(const unsigned char *p, const PcdField &f) {
    switch (f.type) {
        case 'F': if (f.size == 4) { float v; memcpy(&v, p, 4); return v; } else { double v; memcpy(&v, p, 8); return v; }
        case 'U': if (f.size == 1) return *p; if (f.size == 2) { uint16_t v; memcpy(&v, p, 2); return v; }
                  if (f.size == 4) { uint32_t v; memcpy(&v, p, 4); return v; } { uint64_t v; memcpy(&v, p, 8); return (double)v; }
        default:  if (f.size == 1) return *(const signed char *)p; if (f.size == 2) { int16_t v; memcpy(&v, p, 2); return v; }
                  if (f.size == 4) { int32_t v; memcpy(&v, p, 4); return v; } { int64_t v; memcpy(&v, p, 8); return (double)v; }
    }
}

// LZF decompression (the codec of PCD "binary_compressed"): a control byte < 32 starts a literal run of ctrl + 1
// bytes; otherwise the top three bits are the match length - 2 (7 = extended by the next byte) and the low five
// bits plus the following byte the distance - 1 of a back reference into the output.  Returns the number of bytes
// produced, 0 on a malformed stream.
size_t lzf_decompress(const unsigned char *in, size_t in_len, unsigned char *out, size_t out_len) {
    size_t ip = 0, op = 0;
    while (ip < in_len) {
        unsigned ctrl = in[ip++];
        if (ctrl < 32) {
            const size_t run = ctrl + 1;
            if (ip + run > in_len || op + run > out_len) return 0;
            memcpy(out + op, in + ip, run);
            ip += run; op += run;
        } else {
            size_t len = ctrl >> 5;
            if (len == 7) { if (ip >= in_len) return 0; len += in[ip++]; }
            if (ip >= in_len) return 0;
            const size_t dist = ((size_t)(ctrl & 0x1f) << 8) + in[ip++] + 1;
            len += 2;
            if (dist > op || op + len > out_len) return 0;
            for (size_t k = 0; k < len; ++k, ++op) out[op] = out[op - dist];      // the ranges may overlap
        }
    }
    return op;
}
}  // namespace

// Read a PCD file into a malloc'ed packed float4 {x,y,z,intensity} array (*xyzi, release with b2_pcd_free).
static int pcd_read_impl(const char *path, float **xyzi, size_t *n_points) {
    if (!path || !xyzi || !n_points) { set_error("b2_pcd_read: NULL argument"); return B2_ERR_INVALID; }
    *xyzi = nullptr; *n_points = 0;
    FILE *f = fopen(path, "rb");
    if (!f) { set_error("b2_pcd_read: cannot open %s: %s", path, strerror(errno)); return B2_ERR_INVALID; }
    struct Closer { FILE *f; ~Closer() { fclose(f); } } closer{f};      // every return below closes the file
    std::vector<PcdField> fields;
    size_t width = 0, height = 1, points = 0;
    bool have_points = false;
    std::string data;
    char line[4096];
    while (fgets(line, sizeof line, f)) {
        if (line[0] == '#') continue;
        std::vector<std::string> tok;
        for (char *t = strtok(line, " \t\r\n"); t; t = strtok(nullptr, " \t\r\n")) tok.push_back(t);
        if (tok.empty()) continue;
        const std::string &k = tok[0];
        if (k == "FIELDS" || k == "COLUMNS") { fields.assign(tok.size() - 1, PcdField()); for (size_t i = 1; i < tok.size(); ++i) fields[i - 1].name = tok[i]; }
        else if (k == "SIZE")  { for (size_t i = 1; i < tok.size() && i - 1 < fields.size(); ++i) fields[i - 1].size = atoi(tok[i].c_str()); }
        else if (k == "TYPE")  { for (size_t i = 1; i < tok.size() && i - 1 < fields.size(); ++i) fields[i - 1].type = tok[i][0]; }
        else if (k == "COUNT") { for (size_t i = 1; i < tok.size() && i - 1 < fields.size(); ++i) fields[i - 1].count = atoi(tok[i].c_str()); }
        else if (k == "WIDTH"  && tok.size() > 1) width = strtoull(tok[1].c_str(), nullptr, 10);
        else if (k == "HEIGHT" && tok.size() > 1) height = strtoull(tok[1].c_str(), nullptr, 10);
        else if (k == "POINTS" && tok.size() > 1) { points = strtoull(tok[1].c_str(), nullptr, 10); have_points = true; }
        else if (k == "DATA"   && tok.size() > 1) { data = tok[1]; break; }
    }
    if (data.empty() || fields.empty()) { set_error("b2_pcd_read: %s has no PCD header (FIELDS / DATA)", path); return B2_ERR_INVALID; }
    if (!have_points) points = width * height;
    size_t stride = 0;
    int fx = -1, fy = -1, fz = -1, fi = -1;
    for (size_t i = 0; i < fields.size(); ++i) {
        PcdField &fd = fields[i];
        if (fd.size != 1 && fd.size != 2 && fd.size != 4 && fd.size != 8) { set_error("b2_pcd_read: bad SIZE %d", fd.size); return B2_ERR_INVALID; }
        if (fd.type != 'F' && fd.type != 'U' && fd.type != 'I') { set_error("b2_pcd_read: bad TYPE %c", fd.type); return B2_ERR_INVALID; }
        if (fd.type == 'F' && fd.size != 4 && fd.size != 8) { set_error("b2_pcd_read: TYPE F needs SIZE 4 or 8, not %d", fd.size); return B2_ERR_INVALID; }
        if (fd.count < 1) fd.count = 1;
        fd.offset = stride;
        stride += (size_t)fd.size * (size_t)fd.count;
        if (fd.name == "x") fx = (int)i; else if (fd.name == "y") fy = (int)i; else if (fd.name == "z") fz = (int)i;
        else if (fd.name == "intensity") fi = (int)i;
    }
    if (fx < 0 || fy < 0 || fz < 0) { set_error("b2_pcd_read: %s has no x / y / z fields", path); return B2_ERR_INVALID; }
    if (points >= 0xFFFFFFF0ull) { set_error("b2_pcd_read: cloud too large"); return B2_ERR_INVALID; }
    if (stride > (1u << 20)) { set_error("b2_pcd_read: %s: %zu bytes per point is not a point cloud", path, stride); return B2_ERR_INVALID; }
    {
        // the header must not promise more than the file holds (before anything is allocated for it): binary needs
        // stride bytes per point, ascii at least two characters per row, a compressed byte expands at most 88-fold
        const long here = ftell(f);
        long end = here;
        if (here >= 0 && fseek(f, 0, SEEK_END) == 0) { end = ftell(f); fseek(f, here, SEEK_SET); }
        const size_t left = (here >= 0 && end >= here) ? (size_t)(end - here) : 0;
        const bool fits = data == "binary" ? points <= left / (stride ? stride : 1)
                        : data == "ascii" ? points <= left / 2 + 1
                        : data == "binary_compressed" ? points <= (left * 88) / (stride ? stride : 1) + 1 : true;
        if (!fits) { set_error("b2_pcd_read: %s is truncated (%zu points announced, %zu data bytes present)", path, points, left); return B2_ERR_INVALID; }
    }
    float *out = (float *)malloc(points ? points * 16 : 16);
    if (!out) { set_error("b2_pcd_read: out of memory (%zu points)", points); return B2_ERR_INVALID; }
    struct Owner { float *p; ~Owner() { free(p); } } owner{out};          // released on every error path below
    size_t got = 0;
    if (data == "binary") {
        bool direct = fields.size() == 4 && fx == 0 && fy == 1 && fz == 2 && fi == 3 && stride == 16;
        for (size_t i = 0; direct && i < 4; ++i) direct = fields[i].type == 'F' && fields[i].size == 4 && fields[i].count == 1;
        if (direct) {
            got = fread(out, 16, points, f);           // pcl::PointXYZI on disk == the device layout
        } else {
            std::vector<unsigned char> rec(stride);
            for (; got < points && fread(rec.data(), stride, 1, f) == 1; ++got) {
                out[4 * got + 0] = (float)field_value(rec.data() + fields[fx].offset, fields[fx]);
                out[4 * got + 1] = (float)field_value(rec.data() + fields[fy].offset, fields[fy]);
                out[4 * got + 2] = (float)field_value(rec.data() + fields[fz].offset, fields[fz]);
                out[4 * got + 3] = fi >= 0 ? (float)field_value(rec.data() + fields[fi].offset, fields[fi]) : 0.f;
            }
        }
    } else if (data == "binary_compressed") {
        // pcl::PCDWriter::writeBinaryCompressed: u32 compressed size, u32 uncompressed size, one LZF stream holding
        // the cloud field by field (all x, then all y, ...), each field block points * SIZE * COUNT bytes
        uint32_t hdr[2];
        if (fread(hdr, 4, 2, f) != 2) { set_error("b2_pcd_read: %s: truncated compressed header", path); return B2_ERR_INVALID; }
        const size_t csize = hdr[0], usize = hdr[1];
        {
            const long here = ftell(f);
            long end = here;
            if (here >= 0 && fseek(f, 0, SEEK_END) == 0) { end = ftell(f); fseek(f, here, SEEK_SET); }
            if (here < 0 || end < here || (size_t)(end - here) < csize) { set_error("b2_pcd_read: %s: compressed payload of %zu bytes is truncated", path, csize); return B2_ERR_INVALID; }
        }
        if (usize != stride * points) { set_error("b2_pcd_read: %s: uncompressed size %zu != %zu points x %zu bytes", path, usize, points, stride); return B2_ERR_INVALID; }
        std::vector<unsigned char> comp(csize ? csize : 1), raw(usize ? usize : 1);
        if (fread(comp.data(), 1, csize, f) != csize || lzf_decompress(comp.data(), csize, raw.data(), usize) != usize) {
            set_error("b2_pcd_read: %s: corrupt LZF stream", path); return B2_ERR_INVALID;
        }
        auto block = [&](int fld) { return raw.data() + fields[fld].offset * points; };       // offset = bytes of the fields before it
        const size_t sx = (size_t)fields[fx].size * fields[fx].count, sy = (size_t)fields[fy].size * fields[fy].count,
                     sz = (size_t)fields[fz].size * fields[fz].count, si = fi >= 0 ? (size_t)fields[fi].size * fields[fi].count : 0;
        for (; got < points; ++got) {
            out[4 * got + 0] = (float)field_value(block(fx) + got * sx, fields[fx]);
            out[4 * got + 1] = (float)field_value(block(fy) + got * sy, fields[fy]);
            out[4 * got + 2] = (float)field_value(block(fz) + got * sz, fields[fz]);
            out[4 * got + 3] = fi >= 0 ? (float)field_value(block(fi) + got * si, fields[fi]) : 0.f;
        }
    } else if (data == "ascii") {
        std::string row;
        while (got < points) {
            // one ROW per point whatever its length (a row longer than the buffer must not become two points)
            row.clear();
            bool eof = true;
            while (fgets(line, sizeof line, f)) {
                eof = false;
                row += line;
                if (!row.empty() && row.back() == '\n') break;
            }
            if (eof) break;
            std::vector<char> rowbuf(row.begin(), row.end());
            rowbuf.push_back('\0');
            char *line = rowbuf.data();
            size_t col = 0;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            bool any = false;
            char *t = strtok(line, " \t\r\n");
            for (size_t i = 0; i < fields.size() && t; ++i)
                for (int c = 0; c < fields[i].count && t; ++c, ++col, t = strtok(nullptr, " \t\r\n")) {
                    if (c) continue;
                    const float val = (float)strtod(t, nullptr);        // "nan" parses to NaN as in PCL
                    if ((int)i == fx) v[0] = val; else if ((int)i == fy) v[1] = val; else if ((int)i == fz) v[2] = val;
                    else if ((int)i == fi) v[3] = val;
                    any = true;
                }
            if (!any) continue;
            memcpy(out + 4 * got, v, 16);
            ++got;
        }
    } else {
        set_error("b2_pcd_read: DATA %s is not supported (ascii, binary and binary_compressed are)", data.c_str());
        return B2_ERR_INVALID;
    }
    if (got != points) { set_error("b2_pcd_read: %s is truncated (%zu of %zu points)", path, got, points); return B2_ERR_INVALID; }
    owner.p = nullptr;
    *xyzi = out; *n_points = points;
    return 0;
}

// C entry point: no C++ exception (an allocation a hostile header asks for) crosses the ABI
extern "C" int b2_pcd_read(const char *path, float **xyzi, size_t *n_points) {
    try {
        return pcd_read_impl(path, xyzi, n_points);
    } catch (const std::exception &e) {
        set_error("b2_pcd_read: %s: %s", path ? path : "(null)", e.what());
        return B2_ERR_INVALID;
    }
}

extern "C" void b2_pcd_free(float *xyzi) { free(xyzi); }

// Write packed float4 points as pcl::io::savePCDFileBinary writes a pcl::PointCloud<pcl::PointXYZI> (unorganised).
extern "C" int b2_pcd_write_binary(const char *path, const float *xyzi, size_t n_points) {
    if (!path || (n_points && !xyzi)) { set_error("b2_pcd_write_binary: NULL argument"); return B2_ERR_INVALID; }
    FILE *f = fopen(path, "wb");
    if (!f) { set_error("b2_pcd_write_binary: cannot open %s: %s", path, strerror(errno)); return B2_ERR_INVALID; }
    fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\n"
               "COUNT 1 1 1 1\nWIDTH %zu\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA binary\n", n_points, n_points);
    const size_t w = n_points ? fwrite(xyzi, 16, n_points, f) : 0;
    const int rc = fclose(f);
    if (w != n_points || rc != 0) { set_error("b2_pcd_write_binary: short write to %s", path); return B2_ERR_INVALID; }
    return 0;
}

// loadPCDFile straight into a device cloud / savePCDFileBinary from one
extern "C" int b2cloud_upload(b2cloud *c, const void *pts, size_t n, size_t stride, size_t ioff);
extern "C" int b2cloud_download(b2cloud *c, void *out, size_t capacity, size_t stride, size_t ioff, size_t *n);

extern "C" int b2cloud_load_pcd(b2cloud *c, const char *path) {
    if (!c) { set_error("b2cloud_load_pcd: NULL cloud handle"); return B2_ERR_INVALID; }
    float *pts = nullptr;
    size_t n = 0;
    int rc = b2_pcd_read(path, &pts, &n);
    if (rc) return rc;
    rc = b2cloud_upload(c, pts, n, 16, 12);
    free(pts);
    return rc;
}

extern "C" int b2cloud_save_pcd(b2cloud *c, const char *path) {
    if (!c) { set_error("b2cloud_save_pcd: NULL cloud handle"); return B2_ERR_INVALID; }
    float *host = (float *)malloc(c->n * 16 + 16);
    if (!host) { set_error("b2cloud_save_pcd: out of host memory (%zu points)", c->n); return B2_ERR_INVALID; }
    size_t n = 0;
    int rc = b2cloud_download(c, host, c->n, 16, 12, &n);
    if (!rc) rc = b2_pcd_write_binary(path, host, n);
    free(host);
    return rc;
}
