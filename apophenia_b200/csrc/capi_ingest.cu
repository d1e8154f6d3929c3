// capi_ingest.cu -- numeric text files straight into the resident layout.
//
// The reference reads a delimited file with apop_text_to_data (apop_conversions.m4.c:600-690):
// one thread, one character at a time, one apop_matrix_realloc per row; the result is then
// prepped and (here) pinned.  This path mmaps the file, cuts it at line boundaries into
// segments, lets the host threads parse the segments of one group straight into pinned staging
// while the previous group is crossing PCIe, and tiles each group on the device (ols_shuffle
// included when the outcome is column 0).  The apop_data detour -- a second full copy of the
// data in pageable host memory -- disappears.
//
// Text rules honoured (apop_conversions.m4.c:339-403, "Input text file formatting"): the
// delimiter set (default "|,\t"), '#' comments, blank lines, white space and NULs around fields,
// an optional line of column names and an optional leading row-name field (both skipped: names
// do not reach the device), a missing value (NaN) for an empty field unless the delimiter that
// closes it is a space or a tab, C number syntax through strtod (inf, nan, hex floats; a leading number
// followed by other characters converts, as strtod does) with a fast exact path for plain decimals, NaN for a
// field strtod cannot convert -- what the reference's CODE does (apop_conversions.m4.c:659-666, "trouble
// converting data item ...; writing NaN"); its format page still says text is "taken as zeros" (:363) --
// quoted fields ("..." read verbatim, the marks stripped) and backslash escapes.  Fixed-width layouts
// (.field_ends, parse_a_fixed_line, apop_conversions.m4.c:434-464) through the _fixed entry points.
#include "capi_internal.cuh"
#include "host_pool.h"
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <atomic>
#include <string>

using namespace apb;

namespace {

struct TextRules {
    bool delim[256];
    bool has_row_names;
    const int* field_ends = nullptr;    // fixed-width layout: 1-based position of the last character of each field
    int n_ends = 0;
};
inline bool is_blank(unsigned char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\0' || c == '\v' || c == '\f'; }

// exact powers of ten for the fast path (Clinger): a decimal with <= 15 significant digits and
// |exponent| <= 22 is one correctly rounded multiplication or division of two exact doubles
const double P10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// one field [b, e) (already trimmed, not empty) -> double; NaN when strtod converts nothing (apop_conversions.m4.c:659-666)
double field_value(const char* b, const char* e) {
    const char* p = b;
    bool neg = false;
    if (p < e && (*p == '+' || *p == '-')) { neg = (*p == '-'); p++; }
    unsigned long long mant = 0;
    int digits = 0, exp10 = 0;
    bool any = false;
    while (p < e && *p >= '0' && *p <= '9') { if (digits < 19) { mant = mant * 10 + (unsigned)(*p - '0'); digits += (mant != 0); } else exp10++; p++; any = true; }
    if (p < e && *p == '.') {
        p++;
        while (p < e && *p >= '0' && *p <= '9') { if (digits < 19) { mant = mant * 10 + (unsigned)(*p - '0'); digits += (mant != 0); exp10--; } p++; any = true; }
    }
    if (any && p < e && (*p == 'e' || *p == 'E')) {
        const char* q = p + 1;
        bool eneg = false;
        if (q < e && (*q == '+' || *q == '-')) { eneg = (*q == '-'); q++; }
        if (q < e && *q >= '0' && *q <= '9') {
            int ex = 0;
            while (q < e && *q >= '0' && *q <= '9') { if (ex < 10000) ex = ex * 10 + (*q - '0'); q++; }
            exp10 += eneg ? -ex : ex;
            p = q;
        }
    }
    if (any && p == e && digits <= 15 && exp10 >= -22 && exp10 <= 22) {
        double v = (double)mant;
        v = exp10 < 0 ? v / P10[-exp10] : v * P10[exp10];
        return neg ? -v : v;
    }
    // everything else (long mantissas, big exponents, inf, nan, hex floats, text): strtod on a copy
    char buf[64];
    const size_t len = (size_t)(e - b);
    char* heap = nullptr;
    char* s = buf;
    if (len >= sizeof buf) { heap = (char*)malloc(len + 1); s = heap; }
    memcpy(s, b, len);
    s[len] = 0;
    char* end = nullptr;
    const double v = strtod(s, &end);
    const bool converted = end != s;
    if (heap) free(heap);
    return converted ? v : NAN;     // "trouble converting data item ...; writing NaN"
}

// A line that uses quotes or backslashes (apop_conversions.m4.c:352-358): the character after a backslash is an
// ordinary character even if it is a delimiter, # or "; a field surrounded by "s is read verbatim (delimiters and #
// inside it are text) and the marks are stripped.  One character at a time, as the reference reads every line;
// only lines that need it come here.  (A quoted field cannot span lines: the file is cut at newlines first.)
long parse_line_quoted(const char* p, const char* eol, const TextRules& R, double* out, long max_fields) {
    std::string cur;
    bool in_q = false, had_q = false, any = false;
    long n = 0;
    bool skip_first = R.has_row_names;
    auto emit = [&](int closer) {      // closer: the delimiter that ended the field, or -1 at the end of the line
        size_t b = 0, e = cur.size();
        if (!had_q) {
            while (b < e && is_blank((unsigned char)cur[b])) b++;
            while (e > b && is_blank((unsigned char)cur[e - 1])) e--;
        }
        const bool empty = (b == e) && !had_q;
        const bool closed_by_space = closer == ' ' || closer == '\t';
        if (!(empty && closed_by_space)) {
            if (skip_first) skip_first = false;
            else {
                if (out && n < max_fields) out[n] = (b == e) ? NAN : field_value(cur.data() + b, cur.data() + e);   // "" has length 0: NaN
                n++;
            }
        }
        cur.clear();
        had_q = false;
    };
    int last_closer = -1;
    bool pending = false;              // a field is open (something follows the last delimiter, or nothing was read yet)
    for (const char* c = p; c < eol; c++) {
        const unsigned char ch = (unsigned char)*c;
        if (ch == '\\' && c + 1 < eol) { cur.push_back(c[1]); c++; any = true; pending = true; continue; }
        if (ch == '"') { in_q = !in_q; had_q = true; any = true; pending = true; continue; }
        if (!in_q && ch == '#') break;
        if (!in_q && R.delim[ch]) {
            if (!any && is_blank(ch)) continue;        // leading white space of the line, even when it is a delimiter
            emit(ch);
            last_closer = ch;
            pending = false;
            any = true;
            continue;
        }
        if (!is_blank(ch)) { any = true; }
        cur.push_back((char)ch);
        pending = true;
    }
    if (!any) return 0;                                  // blank or comment-only line
    // the last field: what follows the last delimiter; a line ending on a delimiter has a trailing missing value
    // unless that delimiter was a space or a tab
    bool only_blank = true;
    for (char ch : cur) if (!is_blank((unsigned char)ch)) only_blank = false;
    if (pending && !(only_blank && !had_q)) emit(-1);
    else if (last_closer >= 0 && !(last_closer == ' ' || last_closer == '\t')) { cur.clear(); had_q = false; emit(-1); }
    return n;
}

// A line of a fixed-width file (parse_a_fixed_line, apop_conversions.m4.c:434-464): every character up to the
// newline belongs to a field -- no comments, quotes or delimiters -- field i ends with the character at position
// field_ends[i] (1-based), and what follows the last given end is one more field.  A line without characters is
// blank.  Each field then goes through the same conversion as any other (:655-666): strtod skips leading white
// space, an empty or unconvertible field is NaN.
long parse_line_fixed(const char* p, const char* eol, const TextRules& R, double* out, long max_fields) {
    if (p == eol) return 0;
    long n = 0;
    bool skip_first = R.has_row_names;
    const char* f = p;
    for (int i = 0; f < eol; i++) {
        const char* g = (i < R.n_ends && p + R.field_ends[i] < eol) ? p + R.field_ends[i] : eol;
        if (g <= f) {                                   // ends must increase; a repeated or decreasing one never matches again
            g = eol;
        }
        const char* fb = f;
        const char* fe = g;
        while (fb < fe && is_blank((unsigned char)*fb)) fb++;
        while (fe > fb && is_blank((unsigned char)fe[-1])) fe--;
        if (skip_first) skip_first = false;
        else {
            if (out && n < max_fields) out[n] = (fb == fe) ? NAN : field_value(fb, fe);
            n++;
        }
        f = g;
    }
    return n;
}

// Parse one line [p, eol).  out == NULL: only count the fields.  Returns the number of fields
// (0 for a blank / comment-only line).
long parse_line(const char* p, const char* eol, const TextRules& R, double* out, long max_fields) {
    if (R.field_ends) return parse_line_fixed(p, eol, R, out, max_fields);
    // cut the comment
    const char* end = p;
    while (end < eol && *end != '#') { if (*end == '"' || *end == '\\') return parse_line_quoted(p, eol, R, out, max_fields); end++; }
    while (p < end && is_blank((unsigned char)*p)) p++;
    while (end > p && is_blank((unsigned char)end[-1])) end--;
    if (p == end) return 0;
    long n = 0;
    bool skip_first = R.has_row_names;
    const char* f = p;
    for (;;) {
        const char* g = f;
        while (g < end && !R.delim[(unsigned char)*g]) g++;
        // field [f, g), closed by delimiter *g or by the end of the line
        const char* fb = f;
        const char* fe = g;
        while (fb < fe && is_blank((unsigned char)*fb)) fb++;
        while (fe > fb && is_blank((unsigned char)fe[-1])) fe--;
        const bool empty = fb == fe;
        const bool closed_by_space = g < end && (*g == ' ' || *g == '\t');
        if (!(empty && closed_by_space)) {         // an empty field before a space/tab delimiter is a formatting glitch
            if (skip_first) skip_first = false;
            else {
                if (out) { if (n < max_fields) out[n] = empty ? NAN : field_value(fb, fe); }
                n++;
            }
        }
        if (g >= end) break;
        f = g + 1;
        if (f >= end) {                              // the line ends on a delimiter: a trailing missing value unless it was a space/tab
            if (!(*g == ' ' || *g == '\t')) { if (out && n < max_fields) out[n] = NAN; n++; }
            break;
        }
    }
    return n;
}

struct Mapped {
    const char* base = nullptr;
    size_t size = 0;
    int fd = -1;
    bool open(const char* path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) return false;
        size = (size_t)st.st_size;
        if (!size) { base = ""; return true; }
        void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return false;
        madvise(m, size, MADV_SEQUENTIAL);
        base = (const char*)m;
        return true;
    }
    ~Mapped() {
        if (base && size) munmap((void*)base, size);
        if (fd >= 0) close(fd);
    }
};

inline const char* next_line(const char* p, const char* end) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
    return nl ? nl + 1 : end;
}

struct Plan {
    TextRules rules;
    const char* data_begin = nullptr;   // first byte after the column-name line
    const char* end = nullptr;
    long k = 0;                          // numeric columns per row
    std::vector<const char*> seg;        // segment boundaries (line starts), seg.size() = n_seg + 1
    std::vector<size_t> seg_rows;        // data rows per segment
    size_t n_rows = 0;
};

int n_threads() { return HostPool::get().size(); }

template <class F>
void parallel_for(size_t n, F fn) { HostPool::get().run(n, fn); }

// first pass: rules, header, width, segmentation and the number of data rows of every segment
bool make_plan(const Mapped& M, int has_row_names, int has_col_names, const char* delimiters, const int* field_ends, int n_ends,
               size_t seg_bytes, Plan& P) {
    if (n_ends < 0 || (n_ends > 0 && !field_ends)) return set_error("field_ends: %d ends but no array", n_ends);
    for (int i = 0; i < n_ends; i++) if (field_ends[i] <= 0) return set_error("field_ends[%d] = %d: positions are counted from 1", i, field_ends[i]);
    P.rules.field_ends = n_ends > 0 ? field_ends : nullptr;
    P.rules.n_ends = n_ends > 0 ? n_ends : 0;
    memset(P.rules.delim, 0, sizeof P.rules.delim);
    const char* dl = delimiters ? delimiters : "|,\t";      // apop_opts.input_delimiters default
    for (const char* c = dl; *c; c++) P.rules.delim[(unsigned char)*c] = true;
    P.rules.has_row_names = (has_row_names == 1 || has_row_names == 'y' || has_row_names == 'Y');
    const bool names = (has_col_names == 1 || has_col_names == 'y' || has_col_names == 'Y');
    const char* p = M.base;
    P.end = M.base + M.size;
    // (the name line has no entry above the row names, apop_conversions.m4.c:397-400: it is skipped whole)
    if (names) {
        for (;;) {                                           // the first non-blank line holds the column names
            if (p >= P.end) return set_error("the file has no column-name line");
            const char* nl = next_line(p, P.end);
            const char* eol = (nl > p && nl[-1] == '\n') ? nl - 1 : nl;
            // names may be quoted text: only count them when the line is plain
            bool blank = true;
            if (P.rules.field_ends) blank = (eol == p);      // a fixed-width line is blank only when it has no characters
            else for (const char* c = p; c < eol && *c != '#'; c++) if (!is_blank((unsigned char)*c)) { blank = false; break; }
            p = nl;
            if (!blank) break;
        }
    }
    P.data_begin = p;
    // width: the first data line
    for (const char* q = p; q < P.end;) {
        const char* nl = next_line(q, P.end);
        const char* eol = (nl > q && nl[-1] == '\n') ? nl - 1 : nl;
        const long c = parse_line(q, eol, P.rules, nullptr, 0);
        if (c < 0) return set_error("quoted fields and backslash escapes are not supported by the device ingest path");
        if (c > 0) { P.k = c; break; }
        q = nl;
    }
    if (P.k <= 0) return set_error("no data rows in the file");
    // segments of about seg_bytes, cut at line starts
    P.seg.clear();
    P.seg.push_back(P.data_begin);
    while (P.seg.back() < P.end) {
        const char* want = P.seg.back() + seg_bytes;
        if (want >= P.end) { P.seg.push_back(P.end); break; }
        P.seg.push_back(next_line(want, P.end));
    }
    if (P.seg.size() == 1) P.seg.push_back(P.end);
    const size_t ns = P.seg.size() - 1;
    P.seg_rows.assign(ns, 0);
    std::atomic<int> bad(0);
    parallel_for(ns, [&](size_t s) {
        size_t rows = 0;
        for (const char* q = P.seg[s]; q < P.seg[s + 1];) {
            const char* nl = next_line(q, P.seg[s + 1]);
            const char* eol = (nl > q && nl[-1] == '\n') ? nl - 1 : nl;
            const long c = parse_line(q, eol, P.rules, nullptr, 0);
            if (c < 0) bad = 1;
            else if (c > P.k) bad = 2;     // the reference stops reading at such a row (error 't')
            else if (c > 0) rows++;
            q = nl;
        }
        P.seg_rows[s] = rows;
    });
    if (bad == 1) return set_error("quoted fields and backslash escapes are not supported by the device ingest path");
    if (bad == 2) return set_error("a row has more than %ld elements (the width of the first data row)", P.k);
    P.n_rows = 0;
    for (size_t r : P.seg_rows) P.n_rows += r;
    return true;
}

// parse the data rows of segment s into dst (row-major, k per row; short rows are padded with 0
// like the freshly reallocated matrix rows of the reference)
void parse_segment(const Plan& P, size_t s, double* dst) {
    const long k = P.k;
    for (const char* q = P.seg[s]; q < P.seg[s + 1];) {
        const char* nl = next_line(q, P.seg[s + 1]);
        const char* eol = (nl > q && nl[-1] == '\n') ? nl - 1 : nl;
        for (long j = 0; j < k; j++) dst[j] = 0.0;
        const long c = parse_line(q, eol, P.rules, dst, k);
        if (c > 0) dst += k;
        q = nl;
    }
}

}  // namespace

// Host-only reader with the same rules (what apop_text_to_data returns in ->matrix): *out is a
// malloc'ed row-major n x k buffer the caller frees.  Used by the CPU tests; no device needed.
static int text_to_host_impl(const char* path, int has_row_names, int has_col_names, const char* delimiters,
                             const int* field_ends, int n_ends, double** out, size_t* n_rows, size_t* k) {
    LOCK;
    if (!path || !out || !n_rows || !k) { set_error("apop_b200_text_to_host: NULL argument"); return 1; }
    Mapped M;
    if (!M.open(path)) { set_error("trouble opening %s", path); return 1; }
    Plan P;
    if (!make_plan(M, has_row_names, has_col_names, delimiters, field_ends, n_ends, (size_t)8 << 20, P)) return 1;
    double* buf = (double*)malloc(sizeof(double) * std::max<size_t>(P.n_rows * (size_t)P.k, 1));
    if (!buf) { set_error("out of host memory"); return 1; }
    std::vector<size_t> start(P.seg_rows.size() + 1, 0);
    for (size_t s = 0; s < P.seg_rows.size(); s++) start[s + 1] = start[s] + P.seg_rows[s];
    parallel_for(P.seg_rows.size(), [&](size_t s) { parse_segment(P, s, buf + start[s] * (size_t)P.k); });
    *out = buf; *n_rows = P.n_rows; *k = (size_t)P.k;
    return 0;
}

extern "C" int apop_b200_text_to_host(const char* path, int has_row_names, int has_col_names, const char* delimiters,
                                      double** out, size_t* n_rows, size_t* k) {
    return text_to_host_impl(path, has_row_names, has_col_names, delimiters, nullptr, 0, out, n_rows, k);
}
// the fixed-width form (.field_ends of apop_text_to_data): n_field_ends 1-based end positions
extern "C" int apop_b200_text_to_host_fixed(const char* path, int has_row_names, int has_col_names, const int* field_ends,
                                            int n_field_ends, double** out, size_t* n_rows, size_t* k) {
    if (!field_ends || n_field_ends <= 0) { set_error("apop_b200_text_to_host_fixed: no field ends"); return 1; }
    return text_to_host_impl(path, has_row_names, has_col_names, nullptr, field_ends, n_field_ends, out, n_rows, k);
}

static apop_data* text_to_resident_impl(const char* path, int has_row_names, int has_col_names, const char* delimiters,
                                        const int* field_ends, int n_ends, int outcome_col0) {
    LOCK;
    if (!path) { set_error("apop_b200_text_to_resident: NULL path"); return nullptr; }
    if (!require_ready()) return nullptr;
    Runtime& R = rt();
    cudaSetDevice(R.device);
    Mapped M;
    if (!M.open(path)) { set_error("trouble opening %s", path); return nullptr; }
    Plan P;
    const size_t seg_bytes = (size_t)4 << 20;
    if (!make_plan(M, has_row_names, has_col_names, delimiters, field_ends, n_ends, seg_bytes, P)) return nullptr;
    const int k = (int)P.k;
    const bool col0 = outcome_col0 != 0 && k >= 1;
    const size_t n = P.n_rows;
    Resident* r = new Resident();
    r->n_rows = n; r->k = k; r->has_y = col0; r->has_w = false; r->cols = k + (col0 ? 1 : 0);
    r->n_tiles = (n + TILE_ROWS - 1) / TILE_ROWS;
    r->synthetic = true;                  // device-only: there is no host copy to refresh from
    const size_t need = r->n_tiles * (size_t)r->cols * TILE_ROWS;
    if (need && !cuda_ok(cudaMalloc(&r->tiles, need * sizeof(double)), "cudaMalloc(resident tiles)")) { delete r; return nullptr; }

    // groups of segments sized for the staging buffers; a group is parsed by all host threads
    // while the previous one is in flight; rows that do not fill a tile carry over to the next group
    const size_t ns = P.seg_rows.size();
    const size_t group = (size_t)n_threads() * 2;
    size_t max_rows = 0;
    for (size_t g0 = 0; g0 < ns; g0 += group) {
        size_t rows = 0;
        for (size_t s = g0; s < std::min(ns, g0 + group); s++) rows += P.seg_rows[s];
        max_rows = std::max(max_rows, rows);
    }
    const size_t cap_rows = max_rows + TILE_ROWS;
    double *hst[2] = {nullptr, nullptr}, *dst[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool ok = true;
    for (int b = 0; b < 2 && ok; b++)
        ok = cuda_ok(cudaHostAlloc(&hst[b], cap_rows * (size_t)k * sizeof(double), cudaHostAllocDefault), "cudaHostAlloc(ingest staging)") &&
             cuda_ok(cudaMalloc(&dst[b], cap_rows * (size_t)k * sizeof(double)), "cudaMalloc(ingest staging)") &&
             cuda_ok(cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming), "event");
    size_t carry = 0, done_rows = 0;
    int gi = 0;
    for (size_t g0 = 0; ok && g0 < ns; g0 += group, gi++) {
        const int buf = gi & 1;
        if (gi >= 2) ok = cuda_ok(cudaEventSynchronize(ev[buf]), "ingest staging wait");
        // (the carried rows were copied into this buffer's head when the previous group was sent)
        const size_t g1 = std::min(ns, g0 + group);
        std::vector<size_t> start(g1 - g0 + 1, carry);
        for (size_t s = g0; s < g1; s++) start[s - g0 + 1] = start[s - g0] + P.seg_rows[s];
        parallel_for(g1 - g0, [&](size_t i) { parse_segment(P, g0 + i, hst[buf] + start[i] * (size_t)k); });
        const size_t have = start[g1 - g0];
        const bool last = g1 == ns;
        const size_t send = last ? have : have / TILE_ROWS * TILE_ROWS;
        if (send) {
            ok = ok && cuda_ok(cudaMemcpyAsync(dst[buf], hst[buf], send * (size_t)k * sizeof(double), cudaMemcpyHostToDevice, R.stream), "H2D ingest chunk");
            ok = ok && cuda_ok(tile_pack(dst[buf], nullptr, nullptr, send, k, r->cols,
                                         r->tiles + (done_rows / TILE_ROWS) * (size_t)r->cols * TILE_ROWS, R.stream, col0), "tile_pack");
            ok = ok && cuda_ok(cudaEventRecord(ev[buf], R.stream), "event");
        }
        carry = have - send;
        if (carry) {
            if (gi >= 1) ok = ok && cuda_ok(cudaEventSynchronize(ev[buf ^ 1]), "ingest staging wait");   // the other buffer must have left
            memcpy(hst[buf ^ 1], hst[buf] + send * (size_t)k, carry * (size_t)k * sizeof(double));
        }
        done_rows += send;
    }
    ok = ok && cuda_ok(cudaStreamSynchronize(R.stream), "ingest sync");
    for (int b = 0; b < 2; b++) {
        if (hst[b]) cudaFreeHost(hst[b]);
        if (dst[b]) cudaFree(dst[b]);
        if (ev[b]) cudaEventDestroy(ev[b]);
    }
    if (!ok || done_rows != n) {
        if (ok) set_error("ingest: %zu rows tiled, %zu expected", done_rows, n);
        free_resident(r);
        return nullptr;
    }
    apop_data* d = (apop_data*)calloc(1, sizeof(apop_data));
    d->matrix = (gsl_matrix*)calloc(1, sizeof(gsl_matrix));
    d->matrix->size1 = n; d->matrix->size2 = (size_t)k; d->matrix->tda = (size_t)k;
    if (col0) {
        d->vector = (gsl_vector*)calloc(1, sizeof(gsl_vector));
        d->vector->size = n; d->vector->stride = 1;
    }
    R.reg[d] = r;
    return d;
}

extern "C" apop_data* apop_b200_text_to_resident(const char* path, int has_row_names, int has_col_names,
                                                 const char* delimiters, int outcome_col0) {
    return text_to_resident_impl(path, has_row_names, has_col_names, delimiters, nullptr, 0, outcome_col0);
}
extern "C" apop_data* apop_b200_text_to_resident_fixed(const char* path, int has_row_names, int has_col_names,
                                                       const int* field_ends, int n_field_ends, int outcome_col0) {
    if (!field_ends || n_field_ends <= 0) { set_error("apop_b200_text_to_resident_fixed: no field ends"); return nullptr; }
    return text_to_resident_impl(path, has_row_names, has_col_names, nullptr, field_ends, n_field_ends, outcome_col0);
}
