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
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>

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
extern "C" int jb_molecules_pack(const jb_atom_record *const *molecules, const int64_t *counts, int32_t n_molecules,
                                 int64_t *atom_offset, double *x, double *y, double *z, uint32_t *name,
                                 uint32_t *residue_name, int32_t *residue_number, uint8_t *chain_id2, int32_t *serial)
{
    if (n_molecules < 0 || !atom_offset || (n_molecules > 0 && (!molecules || !counts))) return JB_ERR_INVALID;
    int64_t pos = 0;
    atom_offset[0] = 0;
    for (int32_t m = 0; m < n_molecules; m++) {
        const jb_atom_record *a = molecules[m];
        const int64_t n = counts[m];
        if (n < 0 || (n > 0 && !a)) return JB_ERR_INVALID;
        for (int64_t i = 0; i < n; i++, pos++) {
            x[pos] = a[i].x[0]; y[pos] = a[i].x[1]; z[pos] = a[i].x[2];
            memcpy(&name[pos], a[i].name, 4);
            memcpy(&residue_name[pos], a[i].res_name, 4);
            residue_number[pos] = a[i].res_seq;
            chain_id2[pos] = a[i].chain_id2;
            if (serial) serial[pos] = a[i].serial;
        }
        atom_offset[m + 1] = pos;
    }
    return JB_OK;
}
