/*
 * xm_headers.h -- SAM headers on raw bytes (host code, no device work).
 *
 * Restates get_sam_header (xm.py:36-46), add_pg_tag (xm.py:120-131) and the
 * output plan of process_headers (xm.py:133-174) without Python's text layer:
 * the header is the run of leading lines whose first character is '@'; lines
 * end like the reference's 'rt' files end them (universal newlines: "\n",
 * "\r\n" or a lone "\r"); the bytes must be valid UTF-8 (the reference decodes
 * them); the first record starts where the header ends, as a BYTE offset -- no
 * tell() cookie, no seek.
 */
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "xm_walk.h"        /* utf8_valid */

namespace xm {

enum { HDR_OK = 0, HDR_INDEX = 10, HDR_VALUE = 2, HDR_UNICODE = 4, HDR_MORE = -1 };

/* Leading '@' lines of [p, p + n).  whole: the buffer holds the file to its end.  Returns HDR_OK with the lines
 * (without terminators) and the byte offset of the first record; HDR_MORE when the header may go on past the
 * buffer; HDR_INDEX where the reference's `line[0]` fails: an empty file, a file that ends with its header, an
 * empty line right behind the header lines (xm.py:40-43). */
inline int sam_header_scan(const uint8_t *p, uint64_t n, bool whole, std::vector<std::string> &lines, uint64_t &offset)
{
    lines.clear();
    uint64_t at = 0;
    for (;;) {
        if (at >= n) return whole ? HDR_INDEX : HDR_MORE;              /* readline() == '' : ''[0] */
        uint64_t e = at;
        while (e < n && p[e] != '\n' && p[e] != '\r') ++e;
        if (e == n && !whole) return HDR_MORE;
        if (e < n && p[e] == '\r' && e + 1 == n && !whole) return HDR_MORE;     /* "\r\n" may be split across the buffer's end */
        if (e == at) return HDR_INDEX;                                  /* an empty line: ''[0] */
        if (p[at] != '@') { offset = at; return HDR_OK; }
        lines.emplace_back((const char *)p + at, (size_t)(e - at));
        at = e;
        if (at < n) at += (p[at] == '\r' && at + 1 < n && p[at + 1] == '\n') ? 2 : 1;
    }
}

/* str.split() of Python on UTF-8 bytes: ASCII and Unicode white space separate tokens */
inline std::vector<std::string> py_split(const std::string &s)
{
    std::vector<std::string> tok;
    std::string cur;
    auto flush = [&] { if (!cur.empty()) { tok.push_back(cur); cur.clear(); } };
    for (size_t i = 0; i < s.size();) {
        const unsigned char c = (unsigned char)s[i];
        uint32_t cp = c;
        size_t k = 1;
        if (c >= 0xc2 && c < 0xe0 && i + 1 < s.size()) { cp = ((c & 0x1fu) << 6) | ((unsigned char)s[i + 1] & 0x3fu); k = 2; }
        else if (c >= 0xe0 && c < 0xf0 && i + 2 < s.size()) { cp = ((c & 0x0fu) << 12) | (((unsigned char)s[i + 1] & 0x3fu) << 6) | ((unsigned char)s[i + 2] & 0x3fu); k = 3; }
        else if (c >= 0xf0 && i + 3 < s.size()) { cp = 0x10000; k = 4; }
        const bool ws = (cp >= 0x09 && cp <= 0x0d) || (cp >= 0x1c && cp <= 0x20) || cp == 0x85 || cp == 0xa0 || cp == 0x1680 ||
                        (cp >= 0x2000 && cp <= 0x200a) || cp == 0x2028 || cp == 0x2029 || cp == 0x202f || cp == 0x205f || cp == 0x3000;
        if (ws) flush(); else cur.append(s, i, k);
        i += k;
    }
    flush();
    return tok;
}

/* add_pg_tag + the print of process_headers: header lines, Xenomapper's @PG line (chained with PP: to a trailing
 * @PG line's ID), the @CO comment, each followed by "\n".  HDR_INDEX: empty header, or a trailing @PG line without
 * an ID token (xm.py:124-127). */
inline int render_header(const std::vector<std::string> &lines, const char *comment, const char *version, std::string &out)
{
    out.clear();
    if (lines.empty()) return HDR_INDEX;
    std::string pp;
    const std::string &last = lines.back();
    if (last.compare(0, 3, "@PG") == 0) {
        std::string id;
        bool found = false;
        for (auto &t : py_split(last)) if (t.compare(0, 2, "ID") == 0) { id = t; found = true; break; }
        if (!found) return HDR_INDEX;
        pp = "PP" + id.substr(2) + "\t";
    }
    for (auto &l : lines) { out += l; out += '\n'; }
    out += "@PG\tID:Xenomapper\tPN:Xenomapper\t" + pp + "VN:" + version + "\n";
    if (comment && *comment) { out += "@CO\t"; out += comment; out += '\n'; }
    return HDR_OK;
}

struct HeaderPlan {
    uint64_t record_offset[2] = {0, 0};
    int status[6] = {0, 0, 0, 0, 0, 0};
    std::string text[6];
};

/* both headers of the six outputs (xm.py:151-173): primary header for primary_*, unassigned and unresolved */
inline int plan_headers(const std::vector<std::string> h[2], const char *version, HeaderPlan &plan)
{
    static const char *comment[6] = {"species specific reads", "species specific reads", "species specific multimapping reads",
                                     "species specific multimapping reads", "reads that could not be assigned", "reads that could not be resolved"};
    static const int which[6] = {0, 1, 0, 1, 0, 0};
    for (int b = 0; b < 6; ++b) plan.status[b] = render_header(h[which[b]], comment[b], version, plan.text[b]);
    return HDR_OK;
}

}  // namespace xm
