// ingest.cu -- structure ingest: fixed-column PDB text -> packed atom records (HOST code).
//
// once queries take milliseconds, parsing is the end-to-end bottleneck of the
// reference (PDBMoleculeParser.load, src/pyjess/_io.py:68-103, one Python `Atom.loads` ->
// Atom_parse call per line, vendor/jess/src/Atom.c:32-72 as patched).  jb_pdb_parse does
// the line scan and the field extraction of BOTH layers in one pass over the file image and
// writes `jb_atom_record`s (the reference's struct _Atom, eff:Atom.h:29-47, packed) that the
// host layer wraps without any per-atom Python work.  Field semantics are Atom_parse's:
// C atoi/atof PREFIX parsing from the field's first column, fields beyond the end of the
// line keep their zero value, tokens copied with blanks/NULs turned into '_'.
#include <algorithm>
#include <cctype>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/jess_b200.h"

namespace {

inline bool is_space(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// Atom_copyToken (vendor/jess/src/Atom.c:17-30): strncpy + blanks/NUL -> '_'
inline void copy_token(char *d, const char *s, int n)
{
    int i = 0;
    for (; i < n && s[i]; i++) d[i] = is_space((unsigned char)s[i]) ? '_' : s[i];
    for (; i < n; i++) d[i] = '_';
}

// atoi: (int)strtol(s, NULL, 10) of glibc -- blanks, sign, digits, saturating in long
inline int32_t parse_int(const char *s)
{
    while (is_space((unsigned char)*s)) s++;
    bool neg = false;
    if (*s == '+' || *s == '-') { neg = *s == '-'; s++; }
    if (!(*s >= '0' && *s <= '9')) return 0;
    uint64_t v = 0;
    int digits = 0;
    for (; *s >= '0' && *s <= '9'; s++, digits++) {
        if (digits >= 18) return (int32_t)strtol(s - digits - (neg ? 1 : 0), nullptr, 10);   // let libc saturate
        v = v * 10 + (uint64_t)(*s - '0');
    }
    const int64_t sv = neg ? -(int64_t)v : (int64_t)v;
    return (int32_t)sv;
}

const double kPow10[] = { 1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15 };

// atof = strtod(s, NULL).  Fast path for plain decimals with at most 15 significant digits:
// mantissa and 10^k are exact doubles, so ONE correctly rounded division gives the same
// correctly rounded result as strtod.  Anything else (exponents, hex, inf/nan, long
// mantissas) goes to strtod itself.
inline double parse_double(const char *s)
{
    const char *p = s;
    while (is_space((unsigned char)*p)) p++;
    bool neg = false;
    if (*p == '+' || *p == '-') { neg = *p == '-'; p++; }
    uint64_t mant = 0;
    int digits = 0, frac = 0;
    bool any = false;
    for (; *p >= '0' && *p <= '9'; p++) { mant = mant * 10 + (uint64_t)(*p - '0'); digits += (mant != 0); any = true; }
    if (*p == '.') {
        p++;
        for (; *p >= '0' && *p <= '9'; p++) { mant = mant * 10 + (uint64_t)(*p - '0'); digits += (mant != 0); frac++; any = true; }
    }
    const char c = *p;
    if (!any || digits > 15 || frac > 15 || c == 'e' || c == 'E' || c == 'x' || c == 'X' || c == 'n' || c == 'N' ||
        c == 'i' || c == 'I')
        return strtod(s, nullptr);
    const double v = (double)mant / kPow10[frac];
    return neg ? -v : v;
}

inline bool starts(const char *p, int64_t len, const char *lit, int n) { return len >= n && memcmp(p, lit, (size_t)n) == 0; }

inline bool starts_lower(const char *p, int64_t len, const char *lit, int n)
{
    if (len < n) return false;
    for (int i = 0; i < n; i++)
        if (tolower((unsigned char)p[i]) != lit[i]) return false;
    return true;
}

}  // namespace

extern "C" int jb_pdb_parse(const char *text, int64_t n_bytes, int32_t flags, jb_atom_record *records, int64_t capacity,
                            jb_pdb_info *info)
{
    if (!text || n_bytes < 0 || !info || (capacity > 0 && !records)) return JB_ERR_INVALID;
    memset(info, 0, sizeof *info);
    info->header_offset = -1;
    const bool skip_hetatm = (flags & JB_PDB_SKIP_HETATM) != 0, ignore_endmdl = (flags & JB_PDB_IGNORE_ENDMDL) != 0;
    int64_t n_atoms = 0, pos = 0;
    char buf[128];
    while (pos < n_bytes) {
        // one line: up to '\n' (a '\r' before it, or a lone '\r', also ends the line: universal newlines)
        const char *line = text + pos;
        const char *nl = (const char *)memchr(line, '\n', (size_t)(n_bytes - pos));
        int64_t len = nl ? nl - line : n_bytes - pos;
        int64_t next = pos + len + (nl ? 1 : 0);
        const char *cr = (const char *)memchr(line, '\r', (size_t)len);
        if (cr) { len = cr - line; if (!(nl && cr + 1 == nl)) next = pos + len + 1; }
        pos = next;
        const bool atom = starts(line, len, "ATOM", 4);
        if (atom || starts(line, len, "HETATM", 6)) {
            if (!atom && skip_hetatm) continue;
            if (n_atoms >= capacity) { n_atoms++; continue; }          // keep counting: caller may retry
            // Atom.loads appends '\n' + NUL; Atom_parse takes strlen of that (stops at an embedded NUL)
            int n = (int)(len < 120 ? len : 120);
            memcpy(buf, line, (size_t)n);
            buf[n] = '\n'; buf[n + 1] = 0;
            n = (int)strlen(buf);
            jb_atom_record *A = &records[n_atoms++];
            memset(A, 0, sizeof *A);
            const char *s = buf;
            do {
                if (n > 6) A->serial = parse_int(&s[6]); else break;
                if (n > 12) copy_token(A->name, &s[12], 4); else break;
                if (n > 16) A->alt_loc = (uint8_t)s[16]; else break;
                if (n > 17) copy_token(A->res_name, &s[17], 3); else break;
                if (n > 20) A->chain_id1 = (uint8_t)s[20]; else break;
                if (n > 21) A->chain_id2 = is_space((unsigned char)s[21]) ? (uint8_t)'0' : (uint8_t)s[21]; else break;
                if (n > 22) A->res_seq = parse_int(&s[22]); else break;
                if (n > 26) A->i_code = (uint8_t)s[26]; else break;
                if (n > 30) A->x[0] = parse_double(&s[30]); else break;
                if (n > 38) A->x[1] = parse_double(&s[38]); else break;
                if (n > 46) A->x[2] = parse_double(&s[46]); else break;
                if (n > 54) A->occupancy = (float)parse_double(&s[54]); else break;
                if (n > 60) A->temp_factor = (float)parse_double(&s[60]); else break;
                if (n > 72) copy_token(A->seg_id, &s[72], 4); else break;
                if (n > 76) copy_token(A->element, &s[76], 2); else break;
                if (n > 78) A->charge = (int8_t)parse_int(&s[78]);
            } while (0);
        } else if (starts(line, len, "HEADER", 6)) {
            if (info->header_offset < 0) { info->header_offset = line - text; info->header_length = len; }
            info->n_headers++;
        } else if (starts(line, len, "ENDMDL", 6)) {
            if (!ignore_endmdl) break;
        } else if (starts_lower(line, len, "data_", 5) || starts_lower(line, len, "loop_", 5)) {
            info->mmcif_tags = 1;
            break;
        }
    }
    info->n_atoms = n_atoms;
    info->stop_offset = pos;
    return n_atoms > capacity ? JB_ERR_CAPACITY : JB_OK;
}


// ---- mmCIF: the _atom_site loop -> packed atom records (HOST code) ------------------------
// CIFMoleculeParser._load_block (src/pyjess/_io.py:167-219) for well-formed files, in one pass over the
// file image: a CIF 1.1 tokenizer (comments, quoted strings closed by quote + white space, `;` text fields)
// that only looks INTO the tokens of the loop holding `_atom_site.id`, and the reference's row loop (first
// model unless JB_PDB_IGNORE_ENDMDL, ATOM / HETATM selection, running-maximum numbering of HETATM rows
// without a sequence number, `Atom.__init__` encodings).  Everything the reference would answer with an
// exception, and everything Python's int()/float() accept beyond plain decimals, is NOT decided here:
// info->fallback names the reason and the host layer runs its generic loader (pyjess_b200/_cif.py), which
// produces the same records on the files both accept (tests/test_cif.py compares them byte for byte).
namespace {

struct Tok { const char *p; int64_t n; };

struct CifScanner {
    const char *s; int64_t n, i = 0;
    bool bad = false;                                  // unterminated text field
    static bool ws(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r') || (c >= 0x1c && c <= 0x1f); }
    bool line_start(int64_t k) const { return k == 0 || s[k - 1] == '\n' || s[k - 1] == '\r'; }
    bool next(Tok *t)
    {
        for (;;) {
            while (i < n && ws((unsigned char)s[i])) i++;
            if (i >= n) return false;
            const char c = s[i];
            if (c == '#') { while (i < n && s[i] != '\n' && s[i] != '\r') i++; continue; }
            if (c == ';' && line_start(i)) {           // text field: up to the next line starting with ';'
                const int64_t b = i;
                int64_t k = i + 1;
                for (;;) {
                    while (k < n && s[k] != '\n' && s[k] != '\r') k++;
                    if (k >= n) { bad = true; return false; }
                    k++;
                    if (k < n && s[k] == ';') break;
                    if (k >= n) { bad = true; return false; }
                }
                i = k + 1;
                t->p = s + b; t->n = i - b;
                return true;
            }
            const int64_t b = i;
            if (c == '\'' || c == '"') {               // closed by the same quote followed by white space / end
                for (int64_t k = i + 1; k < n && s[k] != '\n' && s[k] != '\r'; k++)
                    if (s[k] == c && (k + 1 >= n || ws((unsigned char)s[k + 1]))) {
                        i = k + 1;
                        t->p = s + b; t->n = i - b;
                        return true;
                    }
            }
            while (i < n && !ws((unsigned char)s[i])) i++;     // bare word (or a quote that never closes)
            t->p = s + b; t->n = i - b;
            return true;
        }
    }
};

inline bool tok_ieq(const Tok &t, const char *lit)
{
    const int64_t m = (int64_t)strlen(lit);
    if (t.n != m) return false;
    for (int64_t k = 0; k < m; k++) if (tolower((unsigned char)t.p[k]) != tolower((unsigned char)lit[k])) return false;
    return true;
}
inline bool tok_iprefix(const Tok &t, const char *lit)
{
    const int64_t m = (int64_t)strlen(lit);
    if (t.n < m) return false;
    for (int64_t k = 0; k < m; k++) if (tolower((unsigned char)t.p[k]) != tolower((unsigned char)lit[k])) return false;
    return true;
}
inline bool tok_eq(const Tok &t, const char *lit) { return t.n == (int64_t)strlen(lit) && memcmp(t.p, lit, (size_t)t.n) == 0; }
inline bool cif_keyword(const Tok &t)
{
    if (t.n == 0) return false;
    if (t.p[0] == '_') return true;
    return tok_iprefix(t, "data_") || tok_iprefix(t, "loop_") || tok_iprefix(t, "save_") || tok_ieq(t, "stop_") || tok_ieq(t, "global_");
}
inline bool tok_ascii(const Tok &t)
{
    for (int64_t k = 0; k < t.n; k++) if ((unsigned char)t.p[k] >= 0x80) return false;
    return true;
}
// int(): plain [+-]digits only (underscores, blanks, other bases: the generic loader decides)
inline bool tok_int(const Tok &t, int64_t *out)
{
    int64_t k = 0;
    bool neg = false;
    if (k < t.n && (t.p[k] == '+' || t.p[k] == '-')) { neg = t.p[k] == '-'; k++; }
    if (k >= t.n || t.n - k > 15) return false;
    int64_t v = 0;
    for (; k < t.n; k++) {
        if (t.p[k] < '0' || t.p[k] > '9') return false;
        v = v * 10 + (t.p[k] - '0');
    }
    *out = neg ? -v : v;
    return true;
}
// float(): [+-]digits[.digits][e[+-]digits] / [+-].digits -- correctly rounded like Python's (strtod)
inline bool tok_float(const Tok &t, double *out)
{
    if (t.n == 0 || t.n > 40) return false;
    int64_t k = 0;
    if (t.p[k] == '+' || t.p[k] == '-') k++;
    int64_t digits = 0;
    while (k < t.n && t.p[k] >= '0' && t.p[k] <= '9') { k++; digits++; }
    if (k < t.n && t.p[k] == '.') {
        k++;
        while (k < t.n && t.p[k] >= '0' && t.p[k] <= '9') { k++; digits++; }
    }
    if (digits == 0) return false;
    if (k < t.n && (t.p[k] == 'e' || t.p[k] == 'E')) {
        k++;
        if (k < t.n && (t.p[k] == '+' || t.p[k] == '-')) k++;
        int64_t ed = 0;
        while (k < t.n && t.p[k] >= '0' && t.p[k] <= '9') { k++; ed++; }
        if (ed == 0) return false;
    }
    if (k != t.n) return false;
    char buf[48];
    memcpy(buf, t.p, (size_t)t.n);
    buf[t.n] = 0;
    *out = parse_double(buf);
    return true;
}
// _encode_token (src/pyjess/_jess.pyx:133-140): n bytes, blank / NUL -> '_'
inline void encode_token(char *d, const char *s, int64_t len, int n)
{
    for (int k = 0; k < n; k++) {
        const char c = k < len ? s[k] : 0;
        d[k] = (c == ' ' || c == 0) ? '_' : c;
    }
}

enum { C_ID, C_ELEM, C_NAME, C_ALT, C_RES, C_CHAIN, C_SEQ, C_ICODE, C_X, C_Y, C_Z, C_OCC, C_B, C_CHARGE, C_GROUP, C_MODEL, C_COUNT };
const char *const kPrimary[C_COUNT] = { "id", "type_symbol", "label_atom_id", "label_alt_id", "label_comp_id", "label_asym_id",
    "label_seq_id", "pdbx_PDB_ins_code", "Cartn_x", "Cartn_y", "Cartn_z", "occupancy", "B_iso_or_equiv", "pdbx_formal_charge",
    "group_PDB", "pdbx_PDB_model_num" };
const char *const kAuthor[C_COUNT] = { "id", "type_symbol", "auth_atom_id", "label_alt_id", "auth_comp_id", "auth_asym_id",
    "auth_seq_id", "pdbx_PDB_ins_code", "Cartn_x", "Cartn_y", "Cartn_z", "occupancy", "B_iso_or_equiv", "pdbx_formal_charge",
    "group_PDB", "pdbx_PDB_model_num" };
inline bool optional_column(int c) { return c == C_ICODE || c == C_CHARGE || c == C_GROUP; }

// the items the host layer reads next to the atom table (id / name / date, _io.py:221-243)
inline bool meta_tag(const Tok &t)
{
    return tok_ieq(t, "_entry.id") || tok_ieq(t, "_struct_keywords.pdbx_keywords") || tok_ieq(t, "_struct.title") ||
           tok_ieq(t, "_pdbx_audit_revision_history.revision_date");
}

}  // namespace

extern "C" int jb_cif_parse(const char *text, int64_t n_bytes, int32_t flags, jb_atom_record *records, int64_t capacity,
                            jb_cif_info *info)
{
    if (!text || !info || n_bytes < 0 || capacity < 0 || (capacity > 0 && !records)) return JB_ERR_INVALID;
    memset(info, 0, sizeof(*info));
    info->loop_offset = info->loop_end = -1;
    const bool skip_hetatm = (flags & JB_PDB_SKIP_HETATM) != 0, ignore_endmdl = (flags & JB_PDB_IGNORE_ENDMDL) != 0;
    const char *const *names = (flags & JB_CIF_USE_AUTHOR) ? kAuthor : kPrimary;
    const int kMaxTags = 256;
    CifScanner sc{ text, n_bytes };
    Tok t;
    bool have = sc.next(&t), found = false;
    int64_t n_atoms = 0;
#define CIF_FALLBACK(code) do { info->fallback = (code); info->n_atoms = n_atoms; return JB_OK; } while (0)
    while (have) {
        if (tok_iprefix(t, "data_")) {
            if (info->n_blocks++ == 0) { info->name_offset = t.p + 5 - text; info->name_length = t.n - 5; }
            have = sc.next(&t);
            continue;
        }
        if (info->n_blocks == 0) CIF_FALLBACK(JB_CIF_SYNTAX);
        if (tok_iprefix(t, "save_")) { have = sc.next(&t); continue; }
        if (!tok_ieq(t, "loop_")) {
            if (t.p[0] != '_') CIF_FALLBACK(JB_CIF_SYNTAX);                   // a value outside any pair / loop
            if (tok_ieq(t, "_atom_site.id")) CIF_FALLBACK(JB_CIF_LAYOUT);     // atom table given as pairs
            const bool meta = meta_tag(t);
            const char *item = t.p;
            have = sc.next(&t);
            if (!have || cif_keyword(t)) CIF_FALLBACK(JB_CIF_SYNTAX);         // tag without a value
            if (meta) {
                if (info->n_meta >= JB_CIF_MAX_META) CIF_FALLBACK(JB_CIF_LAYOUT);
                info->meta[info->n_meta][0] = item - text;
                info->meta[info->n_meta][1] = t.p + t.n - text;
                info->n_meta++;
            }
            have = sc.next(&t);
            continue;
        }
        // ---- a loop: tags, then values up to the next keyword -------------------------------
        const int64_t loop_offset = t.p - text;
        int col[C_COUNT];
        for (int c = 0; c < C_COUNT; c++) col[c] = -1;
        int n_tags = 0;
        bool mine = false, dup = false, meta = false;
        while ((have = sc.next(&t)) && t.p[0] == '_') {
            meta = meta || meta_tag(t);
            if (!found && tok_iprefix(t, "_atom_site.")) {
                const Tok item{ t.p + 11, t.n - 11 };
                if (tok_ieq(item, "id")) mine = true;
                for (int c = 0; c < C_COUNT; c++)
                    if (tok_ieq(item, names[c])) { if (col[c] >= 0) dup = true; col[c] = n_tags; }
            }
            n_tags++;
        }
        if (n_tags == 0) CIF_FALLBACK(JB_CIF_SYNTAX);
        if (!mine) {                                                          // someone else's loop: skip its values
            const char *end = nullptr;
            int64_t n_values = 0;
            for (; have && !cif_keyword(t); have = sc.next(&t), n_values++) end = t.p + t.n;
            if (n_values % n_tags) CIF_FALLBACK(JB_CIF_SYNTAX);                 // ragged loop
            if (meta && end) {
                if (info->n_meta >= JB_CIF_MAX_META) CIF_FALLBACK(JB_CIF_LAYOUT);
                info->meta[info->n_meta][0] = loop_offset;
                info->meta[info->n_meta][1] = end - text;
                info->n_meta++;
            }
            continue;
        }
        found = true;
        info->loop_offset = loop_offset;
        if (dup || n_tags > kMaxTags) CIF_FALLBACK(JB_CIF_LAYOUT);
        for (int c = 0; c < C_COUNT; c++)
            if (col[c] < 0 && !optional_column(c)) CIF_FALLBACK(JB_CIF_LAYOUT);   // "missing columns in CIF files"
        Tok row[kMaxTags];
        bool stopped = false;                  // a later model was met: the remaining rows are only skipped
        int64_t highest = 0;                   // max_residue_number
        int k = 0;
        const char *last_end = t.p;
        for (; have && !cif_keyword(t); have = sc.next(&t)) {
            last_end = t.p + t.n;
            row[k++] = t;
            if (k < n_tags) continue;
            k = 0;
            if (stopped) continue;
            if (!ignore_endmdl && !tok_eq(row[col[C_MODEL]], "1")) { stopped = true; continue; }
            if (col[C_GROUP] < 0) continue;
            const Tok &g = row[col[C_GROUP]];
            const bool het = tok_eq(g, "HETATM");
            if (!tok_eq(g, "ATOM") && (!het || skip_hetatm)) continue;
            for (int c = 0; c < C_COUNT; c++)
                if (col[c] >= 0 && !tok_ascii(row[col[c]])) CIF_FALLBACK(JB_CIF_VALUE);
            int64_t seq = 0, serial = 0, charge = 0;
            const Tok &ts = row[col[C_SEQ]];
            if (het && tok_eq(ts, ".")) { seq = highest; highest++; info->hetatm_unnumbered++; }
            else {
                if (!tok_int(ts, &seq)) CIF_FALLBACK(JB_CIF_VALUE);
                if (seq > highest) highest = seq;
            }
            if (!tok_int(row[col[C_ID]], &serial)) CIF_FALLBACK(JB_CIF_VALUE);
            if (serial < INT32_MIN || serial > INT32_MAX || seq < INT32_MIN || seq > INT32_MAX) CIF_FALLBACK(JB_CIF_VALUE);
            Tok name = row[col[C_NAME]];       // row[2].strip('"')
            while (name.n > 0 && name.p[0] == '"') { name.p++; name.n--; }
            while (name.n > 0 && name.p[name.n - 1] == '"') name.n--;
            const Tok &elem = row[col[C_ELEM]], &alt = row[col[C_ALT]], &res = row[col[C_RES]], &chain = row[col[C_CHAIN]];
            if (name.n > 4 || res.n > 3 || elem.n > 2 || chain.n > 2 || alt.n != 1) CIF_FALLBACK(JB_CIF_VALUE);
            char icode = ' ';
            if (col[C_ICODE] >= 0 && !tok_eq(row[col[C_ICODE]], "?")) {
                if (row[col[C_ICODE]].n != 1) CIF_FALLBACK(JB_CIF_VALUE);
                icode = row[col[C_ICODE]].p[0];
            }
            if (col[C_CHARGE] >= 0 && !tok_eq(row[col[C_CHARGE]], "?")) {
                if (!tok_int(row[col[C_CHARGE]], &charge) || charge < -128 || charge > 127) CIF_FALLBACK(JB_CIF_VALUE);
            }
            double x[3], occ = 0.0, b = 0.0;
            for (int d = 0; d < 3; d++) if (!tok_float(row[col[C_X + d]], &x[d])) CIF_FALLBACK(JB_CIF_VALUE);
            if (!tok_eq(row[col[C_OCC]], ".") && !tok_float(row[col[C_OCC]], &occ)) CIF_FALLBACK(JB_CIF_VALUE);
            if (!tok_eq(row[col[C_B]], ".") && !tok_float(row[col[C_B]], &b)) CIF_FALLBACK(JB_CIF_VALUE);
            if (n_atoms < capacity) {
                jb_atom_record *A = &records[n_atoms];
                memset(A, 0, sizeof(*A));
                A->x[0] = x[0]; A->x[1] = x[1]; A->x[2] = x[2];
                A->occupancy = (float)occ;
                A->temp_factor = (float)b;
                A->serial = (int32_t)serial;
                // Atom.__init__ (pyjess_b200/_model.py, src/pyjess/_jess.pyx:1010-1076): names shorter than four
                // characters get a leading '_'; element right-justified; one-character chains go to chainID2
                char raw[4];
                int len = (int)name.n;
                if (len < 4) { raw[0] = '_'; memcpy(raw + 1, name.p, (size_t)len); len++; } else memcpy(raw, name.p, 4);
                encode_token(A->name, raw, len, 4);
                A->alt_loc = (uint8_t)(tok_eq(alt, ".") ? ' ' : alt.p[0]);
                encode_token(A->res_name, res.p, res.n, 3);
                A->res_name[3] = 0;
                if (chain.n == 2) { A->chain_id1 = (uint8_t)chain.p[0]; A->chain_id2 = (uint8_t)chain.p[1]; }
                else if (chain.n == 1) { A->chain_id1 = ' '; A->chain_id2 = (uint8_t)chain.p[0]; }
                else { A->chain_id1 = ' '; A->chain_id2 = '0'; }
                A->res_seq = (int32_t)seq;
                A->i_code = (uint8_t)icode;
                memcpy(A->seg_id, "____", 4);
                char e2[2] = { 0, 0 };
                if (elem.n == 2) { e2[0] = elem.p[0]; e2[1] = elem.p[1]; } else if (elem.n == 1) e2[1] = elem.p[0];
                encode_token(A->element, e2, 2, 2);
                A->charge = (int8_t)charge;
            }
            n_atoms++;
        }
        if (k != 0) CIF_FALLBACK(JB_CIF_SYNTAX);                              // values do not fill the last row
        info->loop_end = last_end - text;
    }
#undef CIF_FALLBACK
    info->n_atoms = n_atoms;
    if (sc.bad) info->fallback = JB_CIF_SYNTAX;
    else if (info->n_blocks != 1) info->fallback = JB_CIF_BLOCKS;
    else if (!found) info->fallback = JB_CIF_LAYOUT;
    if (info->fallback) return JB_OK;
    return n_atoms > capacity ? JB_ERR_CAPACITY : JB_OK;
}

// ---- output: ATOM lines of a hit ---------------------------------------------------------
// Hit.dump (src/pyjess/_jess.pyx:2541-2621) writes one sprintf'd line per matched atom; this
// is that loop for a whole block of records (decode_token :141-148: '_' and NUL -> ' ').
static inline void decode_token(char *d, const char *s, int n)
{
    for (int i = 0; i < n; i++) d[i] = (s[i] == '_' || s[i] == 0) ? ' ' : s[i];
    d[n] = 0;
}

// "%<width>i": right-justified decimal, wider when it does not fit (like printf)
static inline char *put_int(char *p, int32_t v, int width)
{
    char tmp[12];
    int n = 0;
    uint32_t u = v < 0 ? 0u - (uint32_t)v : (uint32_t)v;
    do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) tmp[n++] = '-';
    for (int i = n; i < width; i++) *p++ = ' ';
    while (n) *p++ = tmp[--n];
    return p;
}

// "%<width>.<prec>f" without the printf machinery.  printf rounds the EXACT binary value half-to-even;
// |v| * 10^prec is computed here with one rounding (<= 2^-23 below 2^30), so the integer nearest to it is
// the right one unless the product sits within 1e-6 of a half: those (exact ties such as k/16, and the
// neighbours no double product can tell apart), non-finite and huge values go through snprintf.
static inline char *put_fixed(char *p, double v, int width, int prec)
{
    const double scale = prec == 3 ? 1000.0 : 100.0;
    const double r = fabs(v) * scale;
    if (!(r < 1073741824.0)) return p + snprintf(p, 400, "%*.*f", width, prec, v);
    const double f = floor(r), frac = r - f;
    if (fabs(frac - 0.5) < 1e-6) return p + snprintf(p, 400, "%*.*f", width, prec, v);
    uint64_t m = (uint64_t)f + (frac > 0.5 ? 1u : 0u);
    char tmp[24];
    int n = 0;
    for (int i = 0; i < prec; i++) { tmp[n++] = (char)('0' + m % 10); m /= 10; }
    tmp[n++] = '.';
    do { tmp[n++] = (char)('0' + m % 10); m /= 10; } while (m);
    if (signbit(v)) tmp[n++] = '-';
    for (int i = n; i < width; i++) *p++ = ' ';
    while (n) *p++ = tmp[--n];
    return p;
}

extern "C" int64_t jb_format_atoms_pdb(const jb_atom_record *atoms, int64_t n, const double *xyz, char *out,
                                       int64_t capacity)
{
    if (!atoms || !out || n < 0) return JB_ERR_INVALID;
    // "ATOM  %5i%5s%c%-3s%c%c%4i%-4c%8.3f%8.3f%8.3f%6.2f%6.2f          %2s\n" (_jess.pyx:1273-1291), field by field;
    // a line is 79 bytes; a double printed with %f takes up to 317, so the field-by-field path wants 2 KiB of room
    char *p = out;
    char name[5], res[4], elem[3];
    for (int64_t k = 0; k < n; k++) {
        const jb_atom_record *a = &atoms[k];
        if (capacity - (p - out) < 2048) {
            if (capacity - (p - out) < 96) return JB_ERR_CAPACITY;
            // short of room for the worst case: let snprintf bound the line
            decode_token(name, a->name, 4);
            decode_token(res, a->res_name, 3);
            decode_token(elem, a->element, 2);
            const double *x = xyz ? xyz + 3 * k : a->x;
            const int64_t room = capacity - (p - out);
            const int w = snprintf(p, (size_t)room,
                                   "ATOM  %5i%5s%c%-3s%c%c%4i%-4c%8.3f%8.3f%8.3f%6.2f%6.2f          %2s\n", a->serial, name,
                                   a->alt_loc ? a->alt_loc : ' ', res, a->chain_id1 ? a->chain_id1 : ' ',
                                   a->chain_id2 ? a->chain_id2 : ' ', a->res_seq, a->i_code ? a->i_code : ' ', x[0], x[1],
                                   x[2], (double)a->occupancy, (double)a->temp_factor, elem);
            if (w < 0 || w >= room) return JB_ERR_CAPACITY;
            p += w;
            continue;
        }
        decode_token(name, a->name, 4);
        decode_token(res, a->res_name, 3);
        decode_token(elem, a->element, 2);
        const double *x = xyz ? xyz + 3 * k : a->x;
        memcpy(p, "ATOM  ", 6); p += 6;
        p = put_int(p, a->serial, 5);
        *p++ = ' ';
        memcpy(p, name, 4); p += 4;
        *p++ = a->alt_loc ? (char)a->alt_loc : ' ';
        memcpy(p, res, 3); p += 3;
        *p++ = a->chain_id1 ? (char)a->chain_id1 : ' ';
        *p++ = a->chain_id2 ? (char)a->chain_id2 : ' ';
        p = put_int(p, a->res_seq, 4);
        *p++ = a->i_code ? (char)a->i_code : ' ';
        memcpy(p, "   ", 3); p += 3;
        p = put_fixed(p, x[0], 8, 3);
        p = put_fixed(p, x[1], 8, 3);
        p = put_fixed(p, x[2], 8, 3);
        p = put_fixed(p, (double)a->occupancy, 6, 2);
        p = put_fixed(p, (double)a->temp_factor, 6, 2);
        memcpy(p, "          ", 10); p += 10;
        memcpy(p, elem, 2); p += 2;
        *p++ = '\n';
    }
    return p - out;
}

// ---- packing: per-molecule _Atom arrays -> the SoA image of jb_molecule_batch -------------
// What Molecule.__init__ + ResIndex_create consume in the reference (SURVEY 8b "bj_molecules_pack"):
// reads the 59-byte records of every molecule once and writes the structure-of-arrays staging
// buffers (pinned or pageable) that jb_molecules_upload / jb_query_host copy to HBM.
static void pack_range(const jb_atom_record *const *molecules, const int64_t *atom_offset, int32_t m0, int32_t m1, double *x,
                       double *y, double *z, uint32_t *name, uint32_t *residue_name, int32_t *residue_number,
                       uint8_t *chain_id2, int32_t *serial)
{
    for (int32_t m = m0; m < m1; m++) {
        const jb_atom_record *a = molecules[m];
        int64_t pos = atom_offset[m];
        const int64_t n = atom_offset[m + 1] - pos;
        for (int64_t i = 0; i < n; i++, pos++) {
            x[pos] = a[i].x[0]; y[pos] = a[i].x[1]; z[pos] = a[i].x[2];
            memcpy(&name[pos], a[i].name, 4);
            memcpy(&residue_name[pos], a[i].res_name, 4);
            residue_number[pos] = a[i].res_seq;
            chain_id2[pos] = a[i].chain_id2;
            if (serial) serial[pos] = a[i].serial;
        }
    }
}

extern "C" int jb_molecules_pack(const jb_atom_record *const *molecules, const int64_t *counts, int32_t n_molecules,
                                 int64_t *atom_offset, double *x, double *y, double *z, uint32_t *name,
                                 uint32_t *residue_name, int32_t *residue_number, uint8_t *chain_id2, int32_t *serial)
{
    if (n_molecules < 0 || !atom_offset || (n_molecules > 0 && (!molecules || !counts))) return JB_ERR_INVALID;
    atom_offset[0] = 0;
    for (int32_t m = 0; m < n_molecules; m++) {
        if (counts[m] < 0 || (counts[m] > 0 && !molecules[m])) return JB_ERR_INVALID;
        atom_offset[m + 1] = atom_offset[m] + counts[m];
    }
    // the copy is a stream of 59-byte reads and 37-41-byte writes (memory bound): large batches are cut into
    // runs of structures with equal shares of the atoms, one host thread each (JB_PACK_THREADS, default <= 8)
    const int64_t total = atom_offset[n_molecules];
    int threads = 1;
    if (total >= 262144 && n_molecules >= 2) {
        const char *env = getenv("JB_PACK_THREADS");
        const int hw = (int)std::thread::hardware_concurrency();
        threads = env ? atoi(env) : std::min(8, hw > 0 ? hw : 1);
        threads = std::max(1, std::min(threads, (int)std::min<int64_t>(n_molecules, total / 131072)));
    }
    if (threads <= 1) {
        pack_range(molecules, atom_offset, 0, n_molecules, x, y, z, name, residue_name, residue_number, chain_id2, serial);
        return JB_OK;
    }
    std::vector<std::thread> pool;
    int32_t m0 = 0;
    for (int t = 0; t < threads; t++) {
        const int64_t want = total * (t + 1) / threads;
        int32_t m1 = t + 1 == threads ? n_molecules
                                      : (int32_t)(std::upper_bound(atom_offset, atom_offset + n_molecules + 1, want) - atom_offset - 1);
        m1 = std::max(m1, m0);
        if (m1 > m0)
            pool.emplace_back(pack_range, molecules, atom_offset, m0, m1, x, y, z, name, residue_name, residue_number, chain_id2,
                              serial);
        m0 = m1;
    }
    for (std::thread &th : pool) th.join();
    return JB_OK;
}
