// Host-side PDB reader for the structure route of predict.py: the per-atom / per-residue tables that
// timed_design_b200/voxelise.py::fast_tables builds with numpy, for MANY files at once on host threads (gzip through zlib).
// The selection rules are fast_tables' (which are parse_pdb's): ATOM records only, states closed by ENDMDL, residues keyed
// by chain + resSeq + iCode in order of first appearance, alternate locations resolved per residue, first occurrence of an
// atom name wins, residue numbers repeated through insertion codes dropped.  Anything unusual (unreadable file, no ATOM
// record, a coordinate field strtod does not consume) is reported per file and the caller re-reads that file with the
// Python parser, so errors and warnings stay the Python ones.  No device work.
#pragma once
#include <zlib.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace tb {

struct PdbResidue {
    char chain;
    char res_id[4];      // stripped, space padded
    char label[3];       // stripped, space padded
    uint8_t has_bb;
    double bb[9];        // N, CA, C
};
struct PdbAtom {
    double xyz[3];
    int32_t name;        // 0 N, 1 CA, 2 C, 3 O, 4 OXT, 5 CB
    int32_t res;         // index into the state's residues
};
struct PdbState {
    int32_t file = 0;
    int32_t n_dup = 0;   // residues dropped because their number repeats (insertion codes)
    std::vector<PdbResidue> res;
    std::vector<PdbAtom> atoms;
};
struct PdbFile {
    int32_t status = 0;  // 0 ok, 1 unreadable, 2 no ATOM records, 3 malformed field
    std::vector<PdbState> states;
};

}  // namespace tb

struct tb_pdb_batch {
    std::vector<tb::PdbFile> files;
};

namespace tb {

static bool pdb_read_file(const char* path, std::string* out) {
    const size_t n = std::strlen(path);
    if (n > 3 && std::strcmp(path + n - 3, ".gz") == 0) {
        gzFile f = gzopen(path, "rb");
        if (!f) return false;
        gzbuffer(f, 1 << 17);
        char buf[1 << 16];
        int got;
        while ((got = gzread(f, buf, sizeof(buf))) > 0) out->append(buf, static_cast<size_t>(got));
        gzclose(f);
        return got == 0;
    }
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    char buf[1 << 16];
    size_t got;
    while ((got = std::fread(buf, 1, sizeof(buf), f)) > 0) out->append(buf, got);
    std::fclose(f);
    return true;
}

// strips leading / trailing blanks of s[0..n) into dst (space padded to cap); returns the stripped length
static int pdb_strip(const char* s, int n, char* dst, int cap) {
    int a = 0, b = n;
    while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r' || s[a] == '\v' || s[a] == '\f')) ++a;
    while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r' || s[b - 1] == '\v' || s[b - 1] == '\f' || s[b - 1] == '\0')) --b;
    int len = b - a;
    for (int i = 0; i < cap; ++i) dst[i] = i < len ? s[a + i] : ' ';
    return len;
}

static bool pdb_field_double(const char* s, int n, double* v) {
    char tmp[16];
    std::memcpy(tmp, s, static_cast<size_t>(n));
    tmp[n] = '\0';
    char* end = nullptr;
    *v = std::strtod(tmp, &end);
    if (end == tmp) return false;
    while (*end == ' ') ++end;
    return *end == '\0';
}

static int pdb_name_code(const char* name4, int len) {
    if (len == 1) return name4[0] == 'N' ? 0 : name4[0] == 'C' ? 2 : name4[0] == 'O' ? 3 : -1;
    if (len == 2) return (name4[0] == 'C' && name4[1] == 'A') ? 1 : (name4[0] == 'C' && name4[1] == 'B') ? 5 : -1;
    if (len == 3) return (name4[0] == 'O' && name4[1] == 'X' && name4[2] == 'T') ? 4 : -1;
    return -1;
}

// one state: `lines` = pointers to ATOM records padded / cut to 54 columns
static bool pdb_parse_state(const std::vector<std::string>& lines, PdbState* st) {
    const size_t n = lines.size();
    struct Rec { uint32_t name; int len; char alt; int res; double xyz[3]; char name4[4]; };
    std::vector<Rec> rec(n);
    std::unordered_map<uint64_t, int> res_index;
    std::vector<size_t> first_atom;
    for (size_t i = 0; i < n; ++i) {
        const char* ln = lines[i].data();
        Rec& r = rec[i];
        r.len = pdb_strip(ln + 12, 4, r.name4, 4);
        std::memcpy(&r.name, r.name4, 4);
        r.alt = ln[16];
        uint64_t key = 0;
        std::memcpy(&key, ln + 21, 6);
        auto it = res_index.find(key);
        if (it == res_index.end()) {
            it = res_index.emplace(key, static_cast<int>(first_atom.size())).first;
            first_atom.push_back(i);
        }
        r.res = it->second;
        if (!pdb_field_double(ln + 30, 8, &r.xyz[0]) || !pdb_field_double(ln + 38, 8, &r.xyz[1]) ||
            !pdb_field_double(ln + 46, 8, &r.xyz[2]))
            return false;
    }
    const int n_res = static_cast<int>(first_atom.size());
    // alternate locations: blank, or the first altLoc seen in the residue (everything when that one is blank);
    // then the first occurrence of an atom name within a residue wins
    std::vector<uint8_t> keep(n);
    std::vector<std::vector<uint32_t>> seen(n_res);
    for (size_t i = 0; i < n; ++i) {
        const Rec& r = rec[i];
        const char res_alt = lines[first_atom[r.res]][16];
        bool k = r.alt == ' ' || r.alt == res_alt || res_alt == ' ';
        if (k) {
            auto& sv = seen[r.res];
            for (uint32_t v : sv)
                if (v == r.name) { k = false; break; }
            if (k) sv.push_back(r.name);
        }
        keep[i] = k;
    }
    // residues, in order of first appearance; a (chain, number) that repeats is dropped
    std::vector<int> remap(n_res, -1);
    std::unordered_map<uint64_t, int> seen_id;
    st->n_dup = 0;
    for (int j = 0; j < n_res; ++j) {
        const char* ln = lines[first_atom[j]].data();
        PdbResidue pr{};
        pr.chain = ln[21] == ' ' ? 'A' : ln[21];
        pdb_strip(ln + 22, 4, pr.res_id, 4);
        pdb_strip(ln + 17, 3, pr.label, 3);
        uint64_t id = 0;
        std::memcpy(&id, &pr.chain, 1);
        std::memcpy(reinterpret_cast<char*>(&id) + 1, pr.res_id, 4);
        if (!seen_id.emplace(id, j).second) { ++st->n_dup; continue; }
        for (int q = 0; q < 9; ++q) pr.bb[q] = 0.0;
        remap[j] = static_cast<int>(st->res.size());
        st->res.push_back(pr);
    }
    std::vector<uint8_t> got(st->res.size() * 3, 0);
    for (size_t i = 0; i < n; ++i) {
        const Rec& r = rec[i];
        if (!keep[i] || remap[r.res] < 0) continue;
        const int code = pdb_name_code(r.name4, r.len);
        if (code < 0) continue;
        const int rj = remap[r.res];
        if (code <= 2) {
            std::memcpy(st->res[rj].bb + 3 * code, r.xyz, sizeof(r.xyz));
            got[rj * 3 + code] = 1;
        }
        PdbAtom a;
        std::memcpy(a.xyz, r.xyz, sizeof(r.xyz));
        a.name = code;
        a.res = rj;
        st->atoms.push_back(a);
    }
    for (size_t j = 0; j < st->res.size(); ++j) st->res[j].has_bb = got[j * 3] && got[j * 3 + 1] && got[j * 3 + 2];
    return true;
}

static void pdb_parse_one(const char* path, int file_index, bool all_states, PdbFile* out) {
    std::string raw;
    if (!pdb_read_file(path, &raw)) { out->status = 1; return; }
    // split into lines; states are closed by ENDMDL
    std::vector<std::vector<std::string>> segs(1);
    size_t pos = 0;
    while (pos <= raw.size()) {
        size_t e = raw.find('\n', pos);
        if (e == std::string::npos) e = raw.size();
        const size_t len = e - pos;
        const char* ln = raw.data() + pos;
        if (len >= 6 && std::memcmp(ln, "ENDMDL", 6) == 0) {
            segs.emplace_back();
        } else if (len >= 6 && std::memcmp(ln, "ATOM  ", 6) == 0) {
            std::string s(ln, std::min<size_t>(len, 54));
            s.resize(54, ' ');
            segs.back().push_back(std::move(s));
        }
        pos = e + 1;
    }
    bool any = false;
    for (auto& sg : segs) {
        if (sg.empty()) continue;
        any = true;
        PdbState st;
        st.file = file_index;
        if (!pdb_parse_state(sg, &st)) { out->status = 3; out->states.clear(); return; }
        out->states.push_back(std::move(st));
        if (!all_states) break;
    }
    if (!any) out->status = 2;
}

}  // namespace tb
