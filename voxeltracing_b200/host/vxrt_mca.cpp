// vxrt_mca.cpp — Minecraft Anvil region reader: the host half of the world import
// (MCWorldImporter::ImportWorld / ImportRegionFile, Core/NBT/Importer.cpp:85-166, which sits on the third-party enkiMI
// reader, Dependencies/enkiMI/enkimi.c).  It only locates and inflates: every chunk section comes out exactly as it is
// stored (4096 block ids in YZX order, 2048 bytes of 4-bit data values, world-space origin); the data-value test, the
// id translation and the scatter into the grid run on the GPU (vxrt_cuda_import_sections, csrc/world.cu).
//
// Format handled: region header of 1024 big-endian (3-byte sector offset, 1-byte sector count) entries
// (enkimi.c:1503-1506, 1946-1966), chunk = 4-byte length + compression type + deflate stream, NBT with the pre-flattening
// chunk layout Level{xPos, zPos, Sections[{Y, Blocks, Data}]} (enkimi.c:2151-2243) — the layout of the engine's
// 'Test MC Worlds' and the only one the engine's 8-bit MC-id table (Core/BlockDatabase.cpp:93-103) is meaningful for.
// Palette sections (1.13+) are counted and skipped.  Behaviour mirrored from the reader the engine uses:
//   * the section index is a signed byte that starts at 0, is set by a "Y" tag when the section has one and advances by
//     one after every section (enkimi.c:2160, 2190-2196, 2232);
//   * a section without "Blocks" contributes nothing; of two sections with the same Y the later one wins as a whole
//     (enkimi.c:2214-2221 overwrites sections[] / dataValues[]);
//   * a chunk needs xPos, zPos and Sections (enkimi.c:2243-2262).
// Divergence (malformed files only): enkiMI matches those tag names at any depth below "Level"; this reader only
// accepts them where the format puts them.
#include "vxrt_host.h"

#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <string>
#include <vector>

namespace {

struct Section {
    bool present = false;
    const uint8_t* blocks = nullptr;
    const uint8_t* data = nullptr;
};

struct Cursor {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;
    bool need(size_t n) {
        if (!ok || (size_t)(end - p) < n) { ok = false; return false; }
        return true;
    }
    uint8_t u8() { if (!need(1)) return 0; return *p++; }
    int32_t be16() { if (!need(2)) return 0; int32_t v = (p[0] << 8) | p[1]; p += 2; return v; }
    int32_t be32() { if (!need(4)) return 0; uint32_t v = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; p += 4; return (int32_t)v; }
    void skip(size_t n) { if (need(n)) p += n; }
};

enum { TAG_END = 0, TAG_BYTE, TAG_SHORT, TAG_INT, TAG_LONG, TAG_FLOAT, TAG_DOUBLE, TAG_BYTE_ARRAY, TAG_STRING, TAG_LIST, TAG_COMPOUND, TAG_INT_ARRAY, TAG_LONG_ARRAY };

void skip_payload(Cursor& c, int type, int depth);

void skip_compound(Cursor& c, int depth) {
    while (c.ok) {
        const int t = c.u8();
        if (t == TAG_END) return;
        c.skip((size_t)c.be16());
        skip_payload(c, t, depth + 1);
    }
}

void skip_payload(Cursor& c, int type, int depth) {
    if (depth > 512) { c.ok = false; return; }
    switch (type) {
        case TAG_BYTE: c.skip(1); break;
        case TAG_SHORT: c.skip(2); break;
        case TAG_INT: case TAG_FLOAT: c.skip(4); break;
        case TAG_LONG: case TAG_DOUBLE: c.skip(8); break;
        case TAG_BYTE_ARRAY: { const int32_t n = c.be32(); if (n < 0) c.ok = false; else c.skip((size_t)n); break; }
        case TAG_STRING: c.skip((size_t)c.be16()); break;
        case TAG_LIST: {
            const int et = c.u8();
            const int32_t n = c.be32();
            for (int32_t i = 0; i < n && c.ok; ++i) skip_payload(c, et, depth + 1);
            break;
        }
        case TAG_COMPOUND: skip_compound(c, depth); break;
        case TAG_INT_ARRAY: { const int32_t n = c.be32(); if (n < 0) c.ok = false; else c.skip((size_t)n * 4); break; }
        case TAG_LONG_ARRAY: { const int32_t n = c.be32(); if (n < 0) c.ok = false; else c.skip((size_t)n * 8); break; }
        default: c.ok = false;
    }
}

bool name_is(const uint8_t* name, int len, const char* s) { return (int)strlen(s) == len && memcmp(name, s, (size_t)len) == 0; }

struct Chunk {
    bool has_x = false, has_z = false, has_sections = false;
    int32_t x = 0, z = 0;
    Section sections[256];
    int palette_sections = 0;
};

// the "Sections" list: elements are compounds {Y, Blocks, Data, ...}
void read_sections(Cursor& c, Chunk& ch) {
    const int et = c.u8();
    const int32_t n = c.be32();
    if (et != TAG_COMPOUND) {  // an empty list is stored with element type End
        for (int32_t i = 0; i < n && c.ok; ++i) skip_payload(c, et, 2);
        return;
    }
    int8_t section_y = 0;
    for (int32_t i = 0; i < n && c.ok; ++i) {
        const uint8_t* blocks = nullptr;
        const uint8_t* data = nullptr;
        bool palette = false;
        while (c.ok) {
            const int t = c.u8();
            if (t == TAG_END) break;
            const int len = c.be16();
            if (!c.need((size_t)len)) break;
            const uint8_t* name = c.p;
            c.p += len;
            if (t == TAG_BYTE_ARRAY && !blocks && name_is(name, len, "Blocks")) {
                const int32_t cnt = c.be32();
                if (cnt >= 4096 && c.need((size_t)cnt)) blocks = c.p;
                c.skip(cnt < 0 ? 0 : (size_t)cnt);
            } else if (t == TAG_BYTE_ARRAY && !data && name_is(name, len, "Data")) {
                const int32_t cnt = c.be32();
                if (cnt >= 2048 && c.need((size_t)cnt)) data = c.p;
                c.skip(cnt < 0 ? 0 : (size_t)cnt);
            } else if (t == TAG_BYTE && name_is(name, len, "Y")) {
                section_y = (int8_t)c.u8();
            } else {
                if (name_is(name, len, "Palette") || name_is(name, len, "BlockStates")) palette = true;
                skip_payload(c, t, 3);
            }
        }
        const int index = (int)section_y + 128;
        if (blocks) {
            ch.sections[index].present = true;
            ch.sections[index].blocks = blocks;
            ch.sections[index].data = data;
        } else if (palette) {
            ch.palette_sections++;
        }
        ++section_y;
    }
}

void read_level(Cursor& c, Chunk& ch) {
    while (c.ok) {
        const int t = c.u8();
        if (t == TAG_END) return;
        const int len = c.be16();
        if (!c.need((size_t)len)) return;
        const uint8_t* name = c.p;
        c.p += len;
        if (t == TAG_INT && !ch.has_x && name_is(name, len, "xPos")) { ch.x = c.be32(); ch.has_x = true; }
        else if (t == TAG_INT && !ch.has_z && name_is(name, len, "zPos")) { ch.z = c.be32(); ch.has_z = true; }
        else if (t == TAG_LIST && !ch.has_sections && name_is(name, len, "Sections")) { ch.has_sections = true; read_sections(c, ch); }
        else skip_payload(c, t, 2);
    }
}

// root: an (unnamed) compound holding "Level"
void read_chunk(const uint8_t* nbt, size_t size, Chunk& ch) {
    Cursor c{nbt, nbt + size};
    if (c.u8() != TAG_COMPOUND) return;
    c.skip((size_t)c.be16());
    while (c.ok) {
        const int t = c.u8();
        if (t == TAG_END) return;
        const int len = c.be16();
        if (!c.need((size_t)len)) return;
        const uint8_t* name = c.p;
        c.p += len;
        if (t == TAG_COMPOUND && name_is(name, len, "Level")) read_level(c, ch);
        else skip_payload(c, t, 1);
    }
}

bool inflate_all(const uint8_t* src, size_t n, int compression, std::vector<uint8_t>& out) {
    out.clear();
    if (compression == 3) { out.assign(src, src + n); return true; }
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, 15 + 32) != Z_OK) return false;  // zlib (type 2) or gzip (type 1) framing
    zs.next_in = const_cast<Bytef*>(src);
    zs.avail_in = (uInt)n;
    // a chunk's NBT is a few hundred KB at most; a stream that inflates past the cap is rejected (counted as a bad chunk)
    // instead of doubling the buffer without bound
    const size_t kMaxInflated = (size_t)16 << 20;
    out.resize(std::min(kMaxInflated, std::max<size_t>(n * 8, 1 << 16)));
    int rc = Z_OK;
    for (;;) {
        zs.next_out = out.data() + zs.total_out;
        zs.avail_out = (uInt)(out.size() - zs.total_out);
        rc = inflate(&zs, Z_NO_FLUSH);
        if (rc == Z_STREAM_END) break;
        if (rc != Z_OK && rc != Z_BUF_ERROR) { inflateEnd(&zs); return false; }
        if (zs.avail_out == 0) {
            if (out.size() >= kMaxInflated) { inflateEnd(&zs); return false; }
            out.resize(std::min(kMaxInflated, out.size() * 2));
        } else if (zs.avail_in == 0) { inflateEnd(&zs); return false; }  // truncated stream
    }
    out.resize(zs.total_out);
    inflateEnd(&zs);
    return true;
}

}  // namespace

struct vxh_mca {
    std::vector<uint8_t> ids, data, has_data;
    std::vector<int32_t> origins;
    int32_t chunks = 0, palette_sections = 0, bad_chunks = 0;
};

extern "C" {

vxh_mca* vxh_mca_open(void) { return new vxh_mca(); }
void vxh_mca_free(vxh_mca* m) { delete m; }

int32_t vxh_mca_add_region_file(vxh_mca* m, const char* path) {
    if (!m || !path) return -1;
    FILE* f = fopen(path, "rb");
    if (!f) return -1;  // the reference throws "Region file not found!" (Importer.cpp:90-94)
    fseek(f, 0, SEEK_END);
    const long size_l = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> file((size_t)(size_l > 0 ? size_l : 0));
    const size_t got = file.empty() ? 0 : fread(file.data(), 1, file.size(), f);
    fclose(f);
    if (got != file.size()) return -2;
    const size_t size = file.size();
    if (size < 8192) return 0;
    int32_t added = 0;
    std::vector<uint8_t> nbt;
    for (int i = 0; i < 1024; ++i) {
        const uint8_t* e = file.data() + 4 * i;
        const size_t loc = (((size_t)e[0] << 16) + ((size_t)e[1] << 8) + e[2]) * 4096;
        if (loc < 8192 || loc + 6 > size) continue;
        const uint8_t* h = file.data() + loc;
        size_t length = ((size_t)h[0] << 24) + ((size_t)h[1] << 16) + ((size_t)h[2] << 8) + h[3];
        const int compression = h[4];
        if (length == 0) continue;
        --length;  // the length counts the compression-type byte
        if (length + loc + 5 > size) continue;
        if (!inflate_all(h + 5, length, compression, nbt) || nbt.empty()) { m->bad_chunks++; continue; }
        Chunk ch;
        read_chunk(nbt.data(), nbt.size(), ch);
        m->palette_sections += ch.palette_sections;
        if (!(ch.has_x && ch.has_z && ch.has_sections)) continue;
        m->chunks++;
        for (int s = 0; s < 256; ++s) {
            const Section& sec = ch.sections[s];
            if (!sec.present) continue;
            m->ids.insert(m->ids.end(), sec.blocks, sec.blocks + 4096);
            if (sec.data) m->data.insert(m->data.end(), sec.data, sec.data + 2048);
            else m->data.insert(m->data.end(), 2048, (uint8_t)0);
            m->has_data.push_back(sec.data ? 1 : 0);
            // enkiGetChunkSectionOrigin (enkimi.c:2282-2289)
            m->origins.push_back(ch.x * 16);
            m->origins.push_back((s - 128) * 16);
            m->origins.push_back(ch.z * 16);
            ++added;
        }
    }
    return added;
}

// every file of the directory whose extension is "mca" (Importer.cpp:152-159); sorted so the batch is deterministic
int32_t vxh_mca_add_region_dir(vxh_mca* m, const char* dir) {
    if (!m || !dir) return -1;
    std::error_code ec;
    std::vector<std::string> files;
    for (auto const& entry : std::filesystem::directory_iterator(dir, ec)) {
        const std::string file = entry.path().string();
        const size_t dot = file.find_last_of('.');
        if (dot != std::string::npos && file.substr(dot + 1) == "mca") files.push_back(file);
    }
    if (ec) return -1;
    std::sort(files.begin(), files.end());
    int32_t total = 0;
    for (const std::string& file : files) {
        const int32_t n = vxh_mca_add_region_file(m, file.c_str());
        if (n < 0) return n;
        total += n;
    }
    return total;
}

int32_t vxh_mca_section_count(const vxh_mca* m) { return m ? (int32_t)m->has_data.size() : 0; }
int32_t vxh_mca_chunk_count(const vxh_mca* m) { return m ? m->chunks : 0; }
int32_t vxh_mca_palette_section_count(const vxh_mca* m) { return m ? m->palette_sections : 0; }
int32_t vxh_mca_bad_chunk_count(const vxh_mca* m) { return m ? m->bad_chunks : 0; }
const uint8_t* vxh_mca_block_ids(const vxh_mca* m) { return m ? m->ids.data() : nullptr; }
const uint8_t* vxh_mca_data_nibbles(const vxh_mca* m) { return m ? m->data.data() : nullptr; }
const uint8_t* vxh_mca_has_data(const vxh_mca* m) { return m ? m->has_data.data() : nullptr; }
const int32_t* vxh_mca_section_origins(const vxh_mca* m) { return m ? m->origins.data() : nullptr; }

}  // extern "C"
