// Mesh ingestion: a threaded reader of Gmsh 2.2 ASCII files with the semantics of the
// reference's Python line loop (jointforces/simulation.py:25-78):
//   * the first line equal to "$Nodes" is followed by the node count and that many lines, of which the tokens
//     after the first are the coordinates (exactly three);
//   * the first line equal to "$Elements" is followed by the element count and that many lines, of which the
//     LAST FOUR tokens are the 1-based node numbers of a tetrahedron (every element line, whatever its type);
//   * an optional "$Jointforces" ... "$EndJointforces" block of key=value lines carries r_inner / r_outer [um].
// Host code only (no device work).  It accelerates the regular case and nothing else: anything the Python loop
// would treat differently from a plain decimal file (carriage returns, non-ASCII bytes, tokens that are not plain
// decimal numbers, short lines, missing sections) is reported as JF_ERR_FORMAT and the caller falls back to the
// line loop, which then raises whatever the reference raises.  Numbers are converted with std::from_chars
// (correctly rounded, locale independent), so coordinates are bit-identical to Python's float().
#include <algorithm>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "jf_internal.h"

struct jf_msh {
    std::vector<char> buf;          // file contents + terminating NUL
    std::vector<size_t> line;       // offsets of line starts, plus one past the end
    int64_t n_nodes = 0, n_elem = 0;
    size_t first_node_line = 0, first_elem_line = 0;
    int has_radii = 0;
    double r_inner = 0.0, r_outer = 0.0;
};

namespace {

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\v' || c == '\f'; }

// tokens of [b, e): at most `cap` (begin, end) pairs are stored, the count of all tokens is returned
inline int tokenize(const char *b, const char *e, const char **tb, const char **te, int cap, bool keep_last) {
    int n = 0;
    while (b < e) {
        while (b < e && is_space(*b)) b++;
        if (b >= e) break;
        const char *s = b;
        while (b < e && !is_space(*b)) b++;
        if (keep_last) {  // ring of the last `cap` tokens
            tb[n % cap] = s;
            te[n % cap] = b;
        } else if (n < cap) {
            tb[n] = s;
            te[n] = b;
        }
        n++;
    }
    return n;
}

inline bool parse_double(const char *s, const char *e, double *out) {
    if (s < e && *s == '+') s++;
    if (s >= e) return false;
    for (const char *p = s; p < e; p++) {
        const char c = *p;
        if (!((c >= '0' && c <= '9') || c == '.' || c == 'e' || c == 'E' || c == '-' || c == '+')) return false;
    }
    auto r = std::from_chars(s, e, *out);
    return r.ec == std::errc() && r.ptr == e;
}

inline bool parse_count(const char *s, const char *e, int64_t *out) {
    while (s < e && is_space(*s)) s++;
    while (e > s && is_space(e[-1])) e--;
    if (s >= e || e - s > 18) return false;
    int64_t v = 0;
    for (const char *p = s; p < e; p++) {
        if (*p < '0' || *p > '9') return false;
        v = v * 10 + (*p - '0');
    }
    *out = v;
    return true;
}

inline bool line_is(const jf_msh *m, size_t i, const char *word) {
    const size_t a = m->line[i], b = m->line[i + 1];
    const size_t n = strlen(word);
    // the reference compares with word + "\n": the newline must be there
    return b - a == n + 1 && m->buf[b - 1] == '\n' && memcmp(&m->buf[a], word, n) == 0;
}

inline const char *line_end(const jf_msh *m, size_t i) {  // end of line i without its newline
    size_t b = m->line[i + 1];
    if (b > m->line[i] && m->buf[b - 1] == '\n') b--;
    return &m->buf[b];
}

int fail(const char *what) {
    jf_set_error("msh reader: %s (fall back to the line loop)", what);
    return JF_ERR_FORMAT;
}

}  // namespace

extern "C" {

int jf_msh_open(const char *path, jf_msh **out, int64_t *n_nodes, int64_t *n_elements, int32_t *has_radii, double *r_inner,
                double *r_outer) {
    if (!path || !out) {
        jf_set_error("jf_msh_open: null argument");
        return JF_ERR_ARG;
    }
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) {
        jf_set_error("jf_msh_open: cannot open %s", path);
        return JF_ERR_IO;
    }
    jf_msh *m = new jf_msh();
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    m->buf.resize((size_t)(sz > 0 ? sz : 0) + 1);
    const size_t got = sz > 0 ? fread(m->buf.data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    if ((long)got != sz) {
        delete m;
        jf_set_error("jf_msh_open: short read of %s", path);
        return JF_ERR_IO;
    }
    m->buf[got] = 0;
    int rc = 0;
    do {
        // universal newlines and text decoding are Python's business: leave such files to the line loop
        if (memchr(m->buf.data(), '\r', got)) { rc = fail("carriage return in file"); break; }
        bool ascii = true;
        {
            const unsigned n_thr = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
            std::vector<int> bad(n_thr, 0);
            std::vector<std::thread> th;
            for (unsigned t = 0; t < n_thr; t++)
                th.emplace_back([&, t] {
                    const size_t a = got * t / n_thr, b = got * (t + 1) / n_thr;
                    const unsigned char *p = (const unsigned char *)m->buf.data();
                    unsigned char acc = 0;
                    for (size_t i = a; i < b; i++) acc |= p[i];
                    // NUL bytes would end a token for the number parser but not for Python
                    bad[t] = (acc & 0x80) || memchr(p + a, 0, b - a) != nullptr;
                });
            for (auto &x : th) x.join();
            for (int b : bad) ascii = ascii && !b;
        }
        if (!ascii) { rc = fail("non-ASCII or NUL byte in file"); break; }
        m->line.reserve(got / 24 + 16);
        for (size_t pos = 0; pos < got;) {
            m->line.push_back(pos);
            const char *nl = (const char *)memchr(&m->buf[pos], '\n', got - pos);
            pos = nl ? (size_t)(nl - m->buf.data()) + 1 : got;
        }
        m->line.push_back(got);
        const size_t n_lines = m->line.size() - 1;
        size_t i_nodes = n_lines, i_elem = n_lines, i_jf = n_lines, i_jfe = n_lines;
        for (size_t i = 0; i < n_lines; i++) {
            if (m->buf[m->line[i]] != '$') continue;
            if (i_nodes == n_lines && line_is(m, i, "$Nodes")) i_nodes = i;
            else if (i_elem == n_lines && line_is(m, i, "$Elements")) i_elem = i;
            else if (i_jf == n_lines && line_is(m, i, "$Jointforces")) i_jf = i;
            else if (i_jfe == n_lines && line_is(m, i, "$EndJointforces")) i_jfe = i;
        }
        if (i_nodes == n_lines || i_elem == n_lines) { rc = fail("no $Nodes or $Elements line"); break; }
        if (i_nodes + 1 >= n_lines || !parse_count(&m->buf[m->line[i_nodes + 1]], line_end(m, i_nodes + 1), &m->n_nodes)) { rc = fail("node count"); break; }
        if (i_elem + 1 >= n_lines || !parse_count(&m->buf[m->line[i_elem + 1]], line_end(m, i_elem + 1), &m->n_elem)) { rc = fail("element count"); break; }
        m->first_node_line = i_nodes + 2;
        m->first_elem_line = i_elem + 2;
        if (m->first_node_line + (size_t)m->n_nodes > n_lines || m->first_elem_line + (size_t)m->n_elem > n_lines) { rc = fail("fewer lines than announced"); break; }
        if ((i_jf == n_lines) != (i_jfe == n_lines)) { rc = fail("unpaired $Jointforces block"); break; }
        if (i_jf != n_lines) {
            if (i_jfe < i_jf) { rc = fail("$EndJointforces before $Jointforces"); break; }
            int seen = 0;
            for (size_t i = i_jf + 1; i < i_jfe && !rc; i++) {
                const char *b = &m->buf[m->line[i]], *e = line_end(m, i);
                while (b < e && is_space(*b)) b++;
                while (e > b && is_space(e[-1])) e--;
                const char *eq = (const char *)memchr(b, '=', (size_t)(e - b));
                if (!eq || memchr(eq + 1, '=', (size_t)(e - eq - 1))) { rc = fail("$Jointforces line is not key=value"); break; }
                const std::string key(b, eq);
                if (key == "r_inner" || key == "r_outer") {
                    const char *vb = eq + 1, *ve = e;
                    while (vb < ve && is_space(*vb)) vb++;
                    double v;
                    if (!parse_double(vb, ve, &v)) { rc = fail("radius is not a plain number"); break; }
                    (key == "r_inner" ? m->r_inner : m->r_outer) = v;  // a repeated key: the last one wins, as in a dict
                    seen |= key == "r_inner" ? 1 : 2;
                }
            }
            if (rc) break;
            if (seen != 3) { rc = fail("$Jointforces block without r_inner and r_outer"); break; }
            m->has_radii = 1;
        }
    } while (0);
    if (rc) {
        delete m;
        return rc;
    }
    if (n_nodes) *n_nodes = m->n_nodes;
    if (n_elements) *n_elements = m->n_elem;
    if (has_radii) *has_radii = m->has_radii;
    if (r_inner) *r_inner = m->r_inner;
    if (r_outer) *r_outer = m->r_outer;
    *out = m;
    return 0;
}

// coords (n_nodes,3) [file units]; tets_f64 (n_elements,4) 0-based as doubles (what the reference returns) and/or
// tets_i32 (n_elements,4) 0-based (what jf_ctx_create takes); either may be NULL.  n_threads <= 0: all host cores.
int jf_msh_read(jf_msh *m, double *coords, double *tets_f64, int32_t *tets_i32, int32_t n_threads) {
    if (!m) {
        jf_set_error("jf_msh_read: null handle");
        return JF_ERR_ARG;
    }
    unsigned n_thr = n_threads > 0 ? (unsigned)n_threads : std::max(1u, std::thread::hardware_concurrency());
    n_thr = std::min(n_thr, 64u);
    std::vector<int> bad(n_thr, 0);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < n_thr; t++)
        th.emplace_back([&, t] {
            const char *tb[4], *te[4];
            if (coords) {
                const int64_t a = m->n_nodes * t / n_thr, b = m->n_nodes * (t + 1) / n_thr;
                for (int64_t i = a; i < b; i++) {
                    const size_t li = m->first_node_line + (size_t)i;
                    if (tokenize(&m->buf[m->line[li]], line_end(m, li), tb, te, 4, false) != 4) { bad[t] = 1; return; }
                    for (int k = 0; k < 3; k++)
                        if (!parse_double(tb[k + 1], te[k + 1], &coords[3 * i + k])) { bad[t] = 1; return; }
                }
            }
            if (tets_f64 || tets_i32) {
                const int64_t a = m->n_elem * t / n_thr, b = m->n_elem * (t + 1) / n_thr;
                for (int64_t i = a; i < b; i++) {
                    const size_t li = m->first_elem_line + (size_t)i;
                    const int n = tokenize(&m->buf[m->line[li]], line_end(m, li), tb, te, 4, true);
                    if (n < 4) { bad[t] = 1; return; }
                    for (int k = 0; k < 4; k++) {
                        const int slot = (n - 4 + k) % 4;
                        int64_t v;
                        if (!parse_count(tb[slot], te[slot], &v) || v > 2147483647ll) { bad[t] = 1; return; }
                        if (tets_f64) tets_f64[4 * i + k] = (double)(v - 1);
                        if (tets_i32) tets_i32[4 * i + k] = (int32_t)(v - 1);
                    }
                }
            }
        });
    for (auto &x : th) x.join();
    for (int b : bad)
        if (b) return fail("a node or element line is not plain decimal");
    return 0;
}

void jf_msh_close(jf_msh *m) { delete m; }

}  // extern "C"
