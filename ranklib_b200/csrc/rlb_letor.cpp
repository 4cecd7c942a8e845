// rlb_letor.cpp — multithreaded reader of LETOR / SVMrank text files into the dense layout rlb_load_dense takes.
//
// Replaces, for the tree-training path, FeatureManager.readInput (R/features/FeatureManager.java:187-245) and
// DataPoint.parse (R/learning/DataPoint.java:58-110); "R/" = src/main/java/ciir/umass/edu/ of the reference.
// Host-only code (no CUDA call): it works on a machine without a GPU.
//
// Semantics kept from the reference, line by line:
//   * a line is trimmed of every char <= ' ' (String.trim); empty lines and lines whose first char is '#' are skipped
//     (FeatureManager.java:199-203); text from the first '#' on is the description and is cut off (DataPoint.java:63-67);
//   * tokens are separated by runs of [ \t\n\v\f\r] (split("\\s+")); token 0 = label (Float.parseFloat, must be >= 0),
//     token 1 = id: the text after its LAST ':' (getValue), tokens 2.. = <fid>:<value> where fid is the text before the
//     FIRST ':' (Integer.parseInt, must be > 0) and value the text after the LAST ':' (Float.parseFloat)
//     (DataPoint.java:44-50,69-83); a repeated fid overwrites; features not listed stay UNKNOWN = NaN, which
//     DenseDataPoint.getFeatureValue reads as 0 (DenseDataPoint.java:21-32);
//   * consecutive lines with the same id form one RankList; a change of id closes the list; with mustHaveRelDoc a list
//     without any label > 0 is dropped (FeatureManager.java:215-231,234-236);
//   * any malformed line is an error (RankLibError in the reference, RLB_E_INVALID + message here).
//
// Design: the file is mapped, cut at line boundaries into one slab per thread, every thread tokenises its slab into a
// compact (fid, value) stream + per-line records; a serial pass stitches the queries together (ids across slab borders)
// and applies the mustHaveRelDoc filter; rlb_letor_fill scatters the streams into the caller's float[N][F] in parallel.
// Decimal -> float conversion is exact (correctly rounded like Float.parseFloat): a Clinger fast path in double whose
// result is accepted only when it is provably not within a double-rounding hazard of a float rounding boundary, else
// strtof (glibc: correctly rounded).
#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/ranklib_b200.h"

const char* rlb_set_error(rlb_ctx* ctx, int code, const char* what, const char* detail);

namespace {

struct Line {
    float label;
    uint32_t id_off, id_len;  // into Slab::ids
    uint64_t feat_off;        // into Slab::fid / Slab::val
    uint32_t nfeat;
};

struct Slab {
    std::vector<Line> lines;
    std::vector<int32_t> fid;
    std::vector<float> val;
    std::string ids;
    int32_t max_fid = 0;
    std::string err;     // first error of the slab
    int64_t err_line = -1;  // slab-relative line ordinal (0-based, counting every physical line)
    int64_t phys_lines = 0;
};

// the class \s of java.util.regex: space, \t, \n, \x0B, \f, \r
inline bool java_space(unsigned char c) { return c <= ' ' && ((1ull << c) & ((1ull << ' ') | (1ull << '\t') | (1ull << '\n') | (1ull << '\v') | (1ull << '\f') | (1ull << '\r'))) != 0; }

// Float.parseFloat on [p, e): optional sign; "NaN"; "Infinity"; decimal or hexadecimal floating literal with an optional
// f/F/d/D suffix; surrounding chars <= ' ' are trimmed by Java.  Returns false on a NumberFormatException.
static const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                  1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// (float)d is the correctly rounded float of the real number x when d = RN_double(x) came from ONE correctly rounded
// operation on exact operands, unless d is exactly the midpoint of two floats (then x may lie on either side of it):
// if a midpoint m lay strictly between x and d, |d - m| <= |d - x| <= ulp(d)/2 with d != m both doubles — impossible.
// Float-subnormal and overflowing magnitudes are left to the slow path.
inline bool double_to_float_is_safe(double d, bool exact) {
    uint64_t bits;
    memcpy(&bits, &d, 8);
    const uint32_t low = (uint32_t)(bits & 0x1fffffffu);
    const int be = (int)((bits >> 52) & 0x7ff);
    return be > 1023 - 126 && be < 1023 + 127 && (exact || low != 0x10000000u);
}

bool parse_java_float(const char* p, const char* e, float* out) {
    {   // the common spelling in LETOR files: [-]digits[.digits] with at most 15 significant digits, nothing else
        const char* q = p;
        const bool neg = (q < e && *q == '-');
        q += neg;
        uint64_t w = 0;
        int nd = 0, frac = 0;
        const char* d0 = q;
        while (q < e && (unsigned)(*q - '0') < 10u) w = w * 10 + (uint64_t)(*q++ - '0');
        nd = (int)(q - d0);
        if (q < e && *q == '.') {
            const char* f0 = ++q;
            while (q < e && (unsigned)(*q - '0') < 10u) w = w * 10 + (uint64_t)(*q++ - '0');
            frac = (int)(q - f0);
        }
        if (q == e && nd + frac > 0 && nd + frac <= 15) {  // <= 15 digits: w < 2^53, no overflow above
            if (w == 0) {
                *out = neg ? -0.0f : 0.0f;
                return true;
            }
            const double d = frac ? (double)w / kPow10[frac] : (double)w;
            if (double_to_float_is_safe(d, frac == 0)) {
                const float f = (float)d;
                *out = neg ? -f : f;
                return true;
            }
        }
    }
    while (p < e && (unsigned char)*p <= ' ') p++;
    while (e > p && (unsigned char)e[-1] <= ' ') e--;
    if (p >= e) return false;
    const char* s = p;
    bool neg = false;
    if (*s == '+' || *s == '-') {
        neg = (*s == '-');
        s++;
    }
    if (s >= e) return false;
    const size_t rem = (size_t)(e - s);
    if (rem == 3 && memcmp(s, "NaN", 3) == 0) {
        *out = std::numeric_limits<float>::quiet_NaN();
        return true;
    }
    if (rem == 8 && memcmp(s, "Infinity", 8) == 0) {
        *out = neg ? -std::numeric_limits<float>::infinity() : std::numeric_limits<float>::infinity();
        return true;
    }
    const bool hex = rem > 2 && s[0] == '0' && (s[1] == 'x' || s[1] == 'X');
    if (!hex) {
        // ---- decimal: digits [. digits] [e[+-]digits] [fFdD] ----
        const char* q = s;
        uint64_t w = 0;
        int nd = 0, dropped = 0, frac = 0;
        bool any = false, inexact_tail = false;
        auto eat = [&](bool after_point) {
            while (q < e && *q >= '0' && *q <= '9') {
                any = true;
                if (nd < 19) {
                    if (w != 0 || *q != '0') {
                        w = w * 10 + (uint64_t)(*q - '0');
                        nd++;
                    }
                    if (after_point) frac++;
                } else {
                    if (*q != '0') inexact_tail = true;
                    if (!after_point) dropped++;
                }
                q++;
            }
        };
        eat(false);
        if (q < e && *q == '.') {
            q++;
            eat(true);
        }
        if (!any) return false;
        long ex = 0;
        if (q < e && (*q == 'e' || *q == 'E')) {
            const char* r = q + 1;
            bool eneg = false;
            if (r < e && (*r == '+' || *r == '-')) {
                eneg = (*r == '-');
                r++;
            }
            if (r >= e || *r < '0' || *r > '9') return false;
            while (r < e && *r >= '0' && *r <= '9') {
                if (ex < 100000) ex = ex * 10 + (*r - '0');
                r++;
            }
            if (eneg) ex = -ex;
            q = r;
        }
        if (q < e && (*q == 'f' || *q == 'F' || *q == 'd' || *q == 'D')) q++;
        if (q != e) return false;
        const long e10 = ex - frac + dropped;
        if (!inexact_tail && w < (1ull << 53) && e10 >= -22 && e10 <= 22) {
            // one correctly rounded double operation on two exact doubles (Clinger)
            const double d = e10 >= 0 ? (double)w * kPow10[e10] : (double)w / kPow10[-e10];
            if (w == 0) {
                *out = neg ? -0.0f : 0.0f;
                return true;
            }
            if (double_to_float_is_safe(d, e10 == 0)) {
                const float f = (float)d;
                *out = neg ? -f : f;
                return true;
            }
        }
    } else {
        // hexadecimal floating literal: validated by strtof below; Java requires the binary exponent
        bool hasp = false;
        for (const char* q = s; q < e; q++) hasp |= (*q == 'p' || *q == 'P');
        if (!hasp) return false;
    }
    // slow path: strtof on a NUL-terminated copy without the Java suffix
    char buf[128];
    size_t n = (size_t)(e - p);
    if (n > 0 && (e[-1] == 'f' || e[-1] == 'F' || ((e[-1] == 'd' || e[-1] == 'D') && !hex))) n--;
    else if (n > 0 && hex && (e[-1] == 'd' || e[-1] == 'D')) {
        // in a hex literal a trailing d/D after the exponent digits is the suffix (exponent digits are decimal)
        n--;
    }
    if (n == 0) return false;
    std::string big;
    const char* z;
    if (n < sizeof(buf)) {
        memcpy(buf, p, n);
        buf[n] = 0;
        z = buf;
    } else {
        big.assign(p, n);
        z = big.c_str();
    }
    // strtof accepts forms Java rejects ("inf", "nan", "infinity" in any case, leading blanks): the decimal grammar
    // was validated above, hex needs its own check of the leading "0x" digits, which strtof does
    char* endp = nullptr;
    errno = 0;
    const float f = strtof(z, &endp);
    if (endp == z || *endp != 0) return false;
    *out = f;
    return true;
}

// Integer.parseInt: optional sign, decimal digits, int range
bool parse_java_int(const char* p, const char* e, int32_t* out) {
    if (p >= e) return false;
    bool neg = false;
    if (*p == '+' || *p == '-') {
        neg = (*p == '-');
        p++;
    }
    if (p >= e) return false;
    int64_t v = 0;
    for (; p < e; p++) {
        if (*p < '0' || *p > '9') return false;
        v = v * 10 + (*p - '0');
        if (v > (int64_t)INT32_MAX + 1) return false;
    }
    if (neg) v = -v;
    if (v > INT32_MAX || v < INT32_MIN) return false;
    *out = (int32_t)v;
    return true;
}

void parse_slab(const char* b, const char* e, Slab& S) {
    const char* p = b;
    int64_t ordinal = -1;
    auto fail = [&](const char* msg) {
        if (S.err.empty()) {
            S.err = msg;
            S.err_line = ordinal;
        }
    };
    while (p < e && S.err.empty()) {
        // BufferedReader.readLine: a line ends at "\n", "\r" or "\r\n"
        const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
        const char* le = nl ? nl : e;
        const char* next = nl ? nl + 1 : e;
        if (const char* cr = (const char*)memchr(p, '\r', (size_t)(le - p))) {
            le = cr;
            next = (cr + 1 < e && cr[1] == '\n') ? cr + 2 : cr + 1;
        }
        ordinal++;
        const char* s = p;
        p = next;
        while (s < le && (unsigned char)*s <= ' ') s++;
        while (le > s && (unsigned char)le[-1] <= ' ') le--;
        if (s >= le || *s == '#') continue;
        if (const char* h = (const char*)memchr(s, '#', (size_t)(le - s))) {
            le = h;
            while (le > s && (unsigned char)le[-1] <= ' ') le--;
        }
        // tokens
        Line L;
        L.feat_off = S.fid.size();
        L.nfeat = 0;
        int tok = 0;
        const char* q = s;
        while (q < le) {
            while (q < le && java_space((unsigned char)*q)) q++;
            if (q >= le) break;
            const char* t0 = q;
            while (q < le && !java_space((unsigned char)*q)) q++;
            const char* t1 = q;
            if (tok == 0) {
                if (!parse_java_float(t0, t1, &L.label)) {
                    fail("Error in DataPoint::parse(): label is not a number");
                    break;
                }
                if (L.label < 0) {
                    fail("Relevance label cannot be negative. System will now exit.");
                    break;
                }
            } else if (tok == 1) {
                const char* c = t1;
                while (c > t0 && c[-1] != ':') c--;  // after the LAST ':' (whole token if there is none)
                L.id_off = (uint32_t)S.ids.size();
                L.id_len = (uint32_t)(t1 - c);
                S.ids.append(c, t1);
            } else {
                // <fid>:<value>; fid = text before the FIRST ':', value = text after the LAST one
                const char* first = t0;
                int32_t f = 0;
                while (first < t1 && (unsigned)(*first - '0') < 10u && f < 100000000) f = f * 10 + (*first++ - '0');
                if (first == t0 || first >= t1 || *first != ':') {  // not plain digits: the general Integer.parseInt
                    first = (const char*)memchr(t0, ':', (size_t)(t1 - t0));
                    if (!first) {
                        fail("Error in DataPoint::parse(): feature token without ':'");
                        break;
                    }
                    if (!parse_java_int(t0, first, &f)) {
                        fail("Error in DataPoint::parse(): feature id is not an integer");
                        break;
                    }
                }
                if (f <= 0) {
                    fail("Cannot use feature numbering less than or equal to zero. Start your features at 1.");
                    break;
                }
                const char* last = t1;
                while (last[-1] != ':') last--;
                float v;
                if (!parse_java_float(last, t1, &v)) {
                    fail("Error in DataPoint::parse(): feature value is not a number");
                    break;
                }
                S.fid.push_back(f);
                S.val.push_back(v);
                L.nfeat++;
                if (f > S.max_fid) S.max_fid = f;
            }
            tok++;
        }
        if (!S.err.empty()) break;
        if (tok < 2) {
            fail("Error in DataPoint::parse(): a line needs a label and a qid");
            break;
        }
        S.lines.push_back(L);
    }
    // count the remaining physical lines for error line numbers of later slabs
    S.phys_lines = ordinal + 1;
    if (!S.err.empty()) return;
}

// gzip input through the system's zlib, bound at first use (no link-time dependency: a plain-text run never needs it)
bool inflate_file(const char* path, std::vector<char>* out, std::string* why) {
    void* z = dlopen("libz.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!z) {
        *why = "cannot read a .gz file: libz.so.1 is not loadable";
        return false;
    }
    typedef void* (*gzopen_t)(const char*, const char*);
    typedef int (*gzread_t)(void*, void*, unsigned);
    typedef int (*gzclose_t)(void*);
    gzopen_t p_open = (gzopen_t)dlsym(z, "gzopen");
    gzread_t p_read = (gzread_t)dlsym(z, "gzread");
    gzclose_t p_close = (gzclose_t)dlsym(z, "gzclose");
    if (!p_open || !p_read || !p_close) {
        *why = "cannot read a .gz file: gzopen/gzread/gzclose not found in libz.so.1";
        return false;
    }
    {   // GZIPInputStream refuses anything without the gzip magic ("Not in GZIP format"); gzread would pass it through
        unsigned char magic[2] = {0, 0};
        FILE* raw = fopen(path, "rb");
        const size_t got = raw ? fread(magic, 1, 2, raw) : 0;
        if (raw) fclose(raw);
        if (got != 2 || magic[0] != 0x1f || magic[1] != 0x8b) {
            *why = std::string("Not in GZIP format: ") + path;
            return false;
        }
    }
    void* f = p_open(path, "rb");
    if (!f) {
        *why = std::string("cannot open ") + path;
        return false;
    }
    std::vector<char> chunk(1u << 22);
    int n;
    while ((n = p_read(f, chunk.data(), (unsigned)chunk.size())) > 0) out->insert(out->end(), chunk.begin(), chunk.begin() + n);
    // a truncated stream decodes up to the cut and then reports Z_BUF_ERROR through gzerror (GZIPInputStream throws
    // "Unexpected end of ZLIB input stream")
    typedef const char* (*gzerror_t)(void*, int*);
    gzerror_t p_error = (gzerror_t)dlsym(z, "gzerror");
    int zerr = 0;
    if (p_error) p_error(f, &zerr);
    p_close(f);
    if (n < 0 || (zerr != 0 && zerr != 1)) {
        *why = std::string("corrupt or truncated gzip stream in ") + path;
        return false;
    }
    return true;
}

}  // namespace

struct rlb_letor {
    std::vector<Slab> slabs;
    // kept data points in file order: (slab, line)
    std::vector<uint32_t> dp_slab;
    std::vector<uint32_t> dp_line;
    std::vector<int32_t> qoff;
    std::vector<std::string> qids;
    int32_t max_fid = 0;
    int64_t entries = 0;  // data points read before the mustHaveRelDoc filter
    int nthreads = 1;
    // a set loaded from the binary cache (rlb_letor_write_binary) has no slabs: rows live here, [n][max_fid], NaN = unknown
    bool dense = false;
    std::vector<float> dense_x, dense_label;
    int64_t n_docs() const { return dense ? (int64_t)dense_label.size() : (int64_t)dp_slab.size(); }
};

namespace {
const char kMagic[16] = "RLB-DENSE-v1\0\0\0";   // 16 bytes

struct BinHeader {
    char magic[16];
    int64_t n_docs, entries;
    int32_t n_lists, max_fid;
    uint64_t qid_bytes;
};

// the binary cache: header | label f32[N] | qoff i32[Q+1] | qids (NUL-terminated, qid_bytes in all) | pad to 64 | X f32[N][max_fid]
int load_binary(const char* path, const char* base, size_t size, bool filter, int nthreads, rlb_letor** out) {
    auto bad = [&](const char* why) {
        rlb_set_error(nullptr, RLB_E_INVALID, "Error in FeatureManager::readInput()", (std::string(why) + " (" + path + ")").c_str());
        return RLB_E_INVALID;
    };
    BinHeader H;
    if (size < sizeof H) return bad("truncated binary cache");
    memcpy(&H, base, sizeof H);
    if (H.n_docs < 0 || H.n_lists < 0 || H.max_fid < 0 || H.n_lists > H.n_docs) return bad("corrupt binary cache header");
    // sizes that cannot fit the file are refused before any product is formed (no overflow in the offsets below)
    if ((uint64_t)H.n_docs > size / 4 || H.qid_bytes > size || (H.max_fid > 0 && (uint64_t)H.n_docs > size / 4 / (uint64_t)H.max_fid))
        return bad("truncated binary cache");
    size_t off = sizeof H;
    const size_t lab_b = (size_t)H.n_docs * 4, qoff_b = ((size_t)H.n_lists + 1) * 4;
    if (size < off + lab_b + qoff_b + H.qid_bytes) return bad("truncated binary cache");
    const float* lab = (const float*)(base + off);
    off += lab_b;
    const int32_t* qoff = (const int32_t*)(base + off);
    off += qoff_b;
    const char* qids = base + off;
    off += (size_t)H.qid_bytes;
    off = (off + 63) & ~(size_t)63;
    const size_t x_b = (size_t)H.n_docs * (size_t)H.max_fid * 4;
    if (size < off + x_b) return bad("truncated binary cache");
    const float* X = (const float*)(base + off);
    if (qoff[0] != 0 || qoff[H.n_lists] != H.n_docs) return bad("corrupt binary cache offsets");
    rlb_letor* h = new rlb_letor();
    h->nthreads = nthreads;
    h->dense = true;
    h->max_fid = H.max_fid;
    h->entries = H.entries;
    h->qoff.push_back(0);
    const char* q = qids;
    const char* qend = qids + H.qid_bytes;
    for (int32_t l = 0; l < H.n_lists; l++) {
        const char* z = (const char*)memchr(q, 0, (size_t)(qend - q));
        if (!z || qoff[l + 1] < qoff[l]) {
            delete h;
            return bad("corrupt binary cache lists");
        }
        bool has_rel = false;
        for (int32_t i = qoff[l]; i < qoff[l + 1]; i++) has_rel |= lab[i] > 0;
        if (!filter || has_rel) {
            h->qids.emplace_back(q, (size_t)(z - q));
            h->dense_label.insert(h->dense_label.end(), lab + qoff[l], lab + qoff[l + 1]);
            h->dense_x.insert(h->dense_x.end(), X + (size_t)qoff[l] * H.max_fid, X + (size_t)qoff[l + 1] * H.max_fid);
            h->qoff.push_back((int32_t)h->dense_label.size());
        }
        q = z + 1;
    }
    *out = h;
    return RLB_OK;
}
}  // namespace

extern "C" {

int rlb_letor_read(const char* path, int32_t must_have_rel_doc, int32_t nthreads, rlb_letor** out) {
    if (!path || !out) return RLB_E_INVALID;
    *out = nullptr;
    int fd = open(path, O_RDONLY);
    if (fd < 0) {
        rlb_set_error(nullptr, RLB_E_INVALID, "Error in FeatureManager::readInput()", strerror(errno));
        return RLB_E_INVALID;
    }
    struct stat sb;
    if (fstat(fd, &sb) != 0) {
        rlb_set_error(nullptr, RLB_E_INVALID, "Error in FeatureManager::readInput()", strerror(errno));
        close(fd);
        return RLB_E_INVALID;
    }
    size_t size = (size_t)sb.st_size;
    const char* base = nullptr;
    std::vector<char> inflated;  // FileUtils.smartReader: a name ending in ".gz" is read through GZIPInputStream
    const size_t plen = strlen(path);
    const bool gz = plen >= 3 && strcmp(path + plen - 3, ".gz") == 0;
    if (gz) {
        close(fd);
        std::string why;
        if (!inflate_file(path, &inflated, &why)) {
            rlb_set_error(nullptr, RLB_E_INVALID, "Error in FeatureManager::readInput()", why.c_str());
            return RLB_E_INVALID;
        }
        size = inflated.size();
        base = inflated.data();
    } else {
        if (size > 0) {
            void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) {
                rlb_set_error(nullptr, RLB_E_NOMEM, "Error in FeatureManager::readInput()", strerror(errno));
                close(fd);
                return RLB_E_NOMEM;
            }
            base = (const char*)m;
            madvise(m, size, MADV_SEQUENTIAL);
        }
        close(fd);
    }
    if (size >= sizeof kMagic && memcmp(base, kMagic, sizeof kMagic) == 0) {  // the binary cache of an earlier read
        if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
        const int rc = load_binary(path, base, size, must_have_rel_doc != 0, nthreads, out);
        if (!gz) munmap((void*)base, size);
        return rc;
    }
    if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    int nslab = (int)std::min<size_t>((size_t)nthreads, std::max<size_t>(1, size / (1u << 12)));
    rlb_letor* h = new rlb_letor();
    h->nthreads = nthreads;
    h->slabs.resize((size_t)nslab);
    // slab borders at line starts
    std::vector<size_t> cut((size_t)nslab + 1, size);
    cut[0] = 0;
    for (int i = 1; i < nslab; i++) {
        size_t c = size / (size_t)nslab * (size_t)i;
        if (c < cut[(size_t)i - 1]) c = cut[(size_t)i - 1];
        const char* nl = c < size ? (const char*)memchr(base + c, '\n', size - c) : nullptr;
        cut[(size_t)i] = nl ? (size_t)(nl - base) + 1 : size;
    }
    {
        std::vector<std::thread> th;
        for (int i = 1; i < nslab; i++)
            th.emplace_back([&, i] { parse_slab(base + cut[(size_t)i], base + cut[(size_t)i + 1], h->slabs[(size_t)i]); });
        parse_slab(base + cut[0], base + cut[1], h->slabs[0]);
        for (auto& t : th) t.join();
    }
    if (base && !gz) munmap((void*)base, size);
    inflated = std::vector<char>();
    // first error in file order
    int64_t line0 = 0;
    for (int i = 0; i < nslab; i++) {
        Slab& S = h->slabs[(size_t)i];
        if (!S.err.empty()) {
            char buf[512];
            snprintf(buf, sizeof buf, "%s (%s, line %lld)", S.err.c_str(), path, (long long)(line0 + S.err_line + 1));
            rlb_set_error(nullptr, RLB_E_INVALID, "Error in FeatureManager::readInput()", buf);
            delete h;
            return RLB_E_INVALID;
        }
        line0 += S.phys_lines;
        h->max_fid = std::max(h->max_fid, S.max_fid);
    }
    // stitch queries: a change of id closes the list (FeatureManager.java:215-221)
    const bool filter = must_have_rel_doc != 0;
    std::string last, first;  // id of the previous line; id of the open list's first line (RankList.getID, RankList.java:68-70)
    bool have_last = false, has_rel = false;
    size_t list_begin = 0;  // index into dp_* where the open list starts
    auto close_list = [&]() {
        if (h->dp_slab.size() == list_begin) return;
        if (!filter || has_rel) {
            h->qids.push_back(first);
            h->qoff.push_back((int32_t)h->dp_slab.size());
            list_begin = h->dp_slab.size();
        } else {
            h->dp_slab.resize(list_begin);
            h->dp_line.resize(list_begin);
        }
    };
    h->qoff.push_back(0);
    for (int i = 0; i < nslab; i++) {
        Slab& S = h->slabs[(size_t)i];
        for (size_t l = 0; l < S.lines.size(); l++) {
            const Line& L = S.lines[l];
            const char* id = S.ids.data() + L.id_off;
            // the reference compares against lastID only when lastID is non-empty (FeatureManager.java:215)
            if (have_last && !last.empty() && (last.size() != L.id_len || memcmp(last.data(), id, L.id_len) != 0)) {
                close_list();
                has_rel = false;
            }
            if (L.label > 0) has_rel = true;
            if (h->dp_slab.size() == list_begin) first.assign(id, L.id_len);
            last.assign(id, L.id_len);
            have_last = true;
            h->dp_slab.push_back((uint32_t)i);
            h->dp_line.push_back((uint32_t)l);
            h->entries++;
        }
    }
    close_list();
    if ((size_t)h->qoff.back() != h->dp_slab.size()) {  // a dropped trailing list
        h->dp_slab.resize((size_t)h->qoff.back());
        h->dp_line.resize((size_t)h->qoff.back());
    }
    *out = h;
    return RLB_OK;
}

int rlb_letor_dims(const rlb_letor* h, int64_t* n_docs, int32_t* n_queries, int32_t* max_fid, int64_t* n_entries) {
    if (!h) return RLB_E_INVALID;
    if (n_docs) *n_docs = h->n_docs();
    if (n_queries) *n_queries = (int32_t)h->qids.size();
    if (max_fid) *max_fid = h->max_fid;
    if (n_entries) *n_entries = h->entries;
    return RLB_OK;
}

int rlb_letor_fill(const rlb_letor* h, const int32_t* feature_ids, int32_t F, float* X, float* label, int32_t* qoff) {
    if (!h || F < 0 || (F > 0 && !feature_ids)) return RLB_E_INVALID;
    const int64_t N = h->n_docs();
    if (qoff) memcpy(qoff, h->qoff.data(), sizeof(int32_t) * h->qoff.size());
    // fid -> column (-1: not selected); a fid listed twice fills both columns
    std::vector<std::vector<int32_t>> dup;
    std::vector<int32_t> col((size_t)h->max_fid + 1, -1);
    bool has_dup = false;
    for (int32_t j = 0; j < F; j++) {
        const int32_t f = feature_ids[j];
        if (f <= 0) {
            rlb_set_error(nullptr, RLB_E_INVALID, "rlb_letor_fill", "feature ids start at 1");
            return RLB_E_INVALID;
        }
        if (f <= h->max_fid) {
            if (col[(size_t)f] >= 0) has_dup = true;
            col[(size_t)f] = j;
        }
    }
    if (has_dup) {
        dup.resize((size_t)h->max_fid + 1);
        for (int32_t j = 0; j < F; j++)
            if (feature_ids[j] <= h->max_fid) dup[(size_t)feature_ids[j]].push_back(j);
    }
    const float unknown = std::numeric_limits<float>::quiet_NaN();
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(h->nthreads, N / 4096 + 1));
    auto work = [&](int t) {
        const int64_t i0 = N * t / nt, i1 = N * (t + 1) / nt;
        for (int64_t i = i0; i < i1 && h->dense; i++) {  // binary cache: a column gather
            if (label) label[i] = h->dense_label[(size_t)i];
            if (!X) continue;
            const float* src = h->dense_x.data() + (size_t)i * (size_t)h->max_fid;
            float* row = X + (size_t)i * (size_t)F;
            for (int32_t j = 0; j < F; j++) row[j] = feature_ids[j] <= h->max_fid ? src[feature_ids[j] - 1] : unknown;
        }
        for (int64_t i = i0; i < i1 && !h->dense; i++) {
            const Slab& S = h->slabs[h->dp_slab[(size_t)i]];
            const Line& L = S.lines[h->dp_line[(size_t)i]];
            if (label) label[i] = L.label;
            if (!X) continue;
            float* row = X + (size_t)i * (size_t)F;
            for (int32_t j = 0; j < F; j++) row[j] = unknown;
            const int32_t* fid = S.fid.data() + L.feat_off;
            const float* val = S.val.data() + L.feat_off;
            for (uint32_t k = 0; k < L.nfeat; k++) {
                const int32_t c = col[(size_t)fid[k]];
                if (c < 0) continue;
                if (!has_dup) row[c] = val[k];
                else
                    for (int32_t cc : dup[(size_t)fid[k]]) row[cc] = val[k];
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    return RLB_OK;
}

int rlb_letor_write_binary(const rlb_letor* h, const char* path) {
    if (!h || !path) return RLB_E_INVALID;
    const int64_t N = h->n_docs();
    const int32_t F = h->max_fid;
    std::vector<int32_t> fid((size_t)F);
    for (int32_t j = 0; j < F; j++) fid[(size_t)j] = j + 1;
    std::vector<float> X((size_t)N * (size_t)F), lab((size_t)N);
    if (int rc = rlb_letor_fill(h, fid.data(), F, X.data(), lab.data(), nullptr)) return rc;
    BinHeader H;
    memset(&H, 0, sizeof H);
    memcpy(H.magic, kMagic, sizeof kMagic);
    H.n_docs = N;
    H.entries = h->entries;
    H.n_lists = (int32_t)h->qids.size();
    H.max_fid = F;
    std::string qids;
    for (const auto& q : h->qids) {
        if (q.find('\0') != std::string::npos) {
            rlb_set_error(nullptr, RLB_E_INVALID, "rlb_letor_write_binary", "a list id contains a NUL byte");
            return RLB_E_INVALID;
        }
        qids.append(q).push_back('\0');
    }
    H.qid_bytes = qids.size();
    FILE* f = fopen(path, "wb");
    if (!f) {
        rlb_set_error(nullptr, RLB_E_INVALID, "rlb_letor_write_binary", strerror(errno));
        return RLB_E_INVALID;
    }
    size_t off = 0;
    bool ok = true;
    auto put = [&](const void* p, size_t n) {
        ok = ok && (n == 0 || fwrite(p, 1, n, f) == n);
        off += n;
    };
    put(&H, sizeof H);
    put(lab.data(), lab.size() * 4);
    put(h->qoff.data(), h->qoff.size() * 4);
    put(qids.data(), qids.size());
    static const char zeros[64] = {0};
    put(zeros, ((off + 63) & ~(size_t)63) - off);
    put(X.data(), X.size() * 4);
    ok = (fclose(f) == 0) && ok;
    if (!ok) {
        rlb_set_error(nullptr, RLB_E_INVALID, "rlb_letor_write_binary", "short write");
        return RLB_E_INVALID;
    }
    return RLB_OK;
}

int rlb_load_letor(rlb_ctx* ctx, const rlb_letor* h, const int32_t* feature_ids, int32_t F) {
    if (!ctx || !h || F < 0) return RLB_E_INVALID;
    std::vector<int32_t> all;
    if (!feature_ids) {  // FeatureManager.getFeatureFromSampleVector: every feature id up to the largest seen
        F = h->max_fid;
        all.resize((size_t)F);
        for (int32_t j = 0; j < F; j++) all[(size_t)j] = j + 1;
        feature_ids = all.data();
    }
    const int64_t N = h->n_docs();
    std::vector<float> X((size_t)N * (size_t)F), label((size_t)N);
    std::vector<int32_t> qoff(h->qoff.size());
    if (int rc = rlb_letor_fill(h, feature_ids, F, X.data(), label.data(), qoff.data())) return rc;
    return rlb_load_dense(ctx, X.data(), N, F, feature_ids, label.data(), qoff.data(), (int32_t)h->qids.size());
}

const char* rlb_letor_qid(const rlb_letor* h, int32_t q) {
    if (!h || q < 0 || (size_t)q >= h->qids.size()) return "";
    return h->qids[(size_t)q].c_str();
}

int rlb_letor_free(rlb_letor* h) {
    delete h;
    return RLB_OK;
}

/* Float.parseFloat for one token (test hook of the reader's number grammar): returns RLB_OK and the float, or
 * RLB_E_INVALID where Java throws NumberFormatException. */
int rlb_parse_java_float(const char* text, float* out) {
    if (!text || !out) return RLB_E_INVALID;
    return parse_java_float(text, text + strlen(text), out) ? RLB_OK : RLB_E_INVALID;
}

}  // extern "C"
