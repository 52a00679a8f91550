// mkhost.cpp — host side above the mksnap C-ABI (see include/mkhost.h).
//
// C++ stand-in for the Go code a cgo build of makisu would keep on this path: the filepath.Walk-ordered
// context stream (cacheID), the MemFS copy-op layer (entry order + tar headers) and the arena packer.
// It does no hashing; digests come from libmksnap (GPU).  Reference line numbers are cited per function.
#include "../../include/mkhost.h"

#include <dirent.h>
#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace {

struct HostError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

std::string errno_str(const std::string &what, const std::string &path)
{
    return what + " " + path + ": " + strerror(errno);
}

// ---------------------------------------------------------------------------------------------------
// Go path helpers (path.Clean / filepath.Join / filepath.Rel, lexical)
// ---------------------------------------------------------------------------------------------------
std::string go_clean(const std::string &p)
{
    if (p.empty())
        return ".";
    const bool rooted = p[0] == '/';
    std::vector<std::string> parts;
    size_t i = 0;
    while (i < p.size()) {
        while (i < p.size() && p[i] == '/')
            ++i;
        size_t j = i;
        while (j < p.size() && p[j] != '/')
            ++j;
        if (j > i) {
            std::string e = p.substr(i, j - i);
            if (e == ".") {
            } else if (e == "..") {
                if (!parts.empty() && parts.back() != "..")
                    parts.pop_back();
                else if (!rooted)
                    parts.push_back("..");
            } else {
                parts.push_back(e);
            }
        }
        i = j;
    }
    std::string out = rooted ? "/" : "";
    for (size_t k = 0; k < parts.size(); ++k) {
        if (k)
            out += "/";
        out += parts[k];
    }
    return out.empty() ? "." : out;
}

std::string go_join(const std::string &a, const std::string &b)
{
    if (a.empty())
        return b.empty() ? "" : go_clean(b);
    if (b.empty())
        return go_clean(a);
    return go_clean(a + "/" + b);
}

std::string go_rel(const std::string &basepath, const std::string &targpath)
{
    const std::string base = go_clean(basepath), targ = go_clean(targpath);
    if (base == targ)
        return ".";
    auto split = [](const std::string &s) {
        std::vector<std::string> v;
        size_t i = 0;
        while (i < s.size()) {
            size_t j = s.find('/', i);
            if (j == std::string::npos)
                j = s.size();
            if (j > i)
                v.push_back(s.substr(i, j - i));
            i = j + 1;
        }
        return v;
    };
    auto bv = split(base == "." ? "" : base), tv = split(targ == "." ? "" : targ);
    size_t k = 0;
    while (k < bv.size() && k < tv.size() && bv[k] == tv[k])
        ++k;
    std::string out;
    for (size_t i = k; i < bv.size(); ++i)
        out += out.empty() ? ".." : "/..";
    for (size_t i = k; i < tv.size(); ++i)
        out += (out.empty() ? "" : "/") + tv[i];
    return out.empty() ? "." : out;
}

// lib/pathutils/path.go:41-68
std::string abs_path(const std::string &p)
{
    std::string t = p;
    while (!t.empty() && t.back() == '/')
        t.pop_back();
    return go_clean("/" + t);
}
std::string rel_path(const std::string &p)
{
    size_t i = 0;
    while (i < p.size() && p[i] == '/')
        ++i;
    return p.substr(i);
}
std::vector<std::string> split_path(const std::string &p)
{
    std::vector<std::string> v;
    size_t i = 0;
    while (i < p.size()) {
        while (i < p.size() && p[i] == '/')
            ++i;
        size_t j = i;
        while (j < p.size() && p[j] != '/')
            ++j;
        if (j > i)
            v.push_back(p.substr(i, j - i));
        i = j;
    }
    return v;
}
std::string path_base(const std::string &p)
{
    std::string t = p;
    while (t.size() > 1 && t.back() == '/')
        t.pop_back();
    size_t i = t.rfind('/');
    return i == std::string::npos ? t : t.substr(i + 1);
}

// ---------------------------------------------------------------------------------------------------
// filepath.Walk / Match / Glob (go1.14)
// ---------------------------------------------------------------------------------------------------
std::vector<std::string> sorted_names(const std::string &dir)
{
    std::vector<std::string> names;
    DIR *d = opendir(dir.c_str());
    if (!d)
        throw HostError(errno_str("open", dir));
    while (struct dirent *e = readdir(d)) {
        if (!strcmp(e->d_name, ".") || !strcmp(e->d_name, ".."))
            continue;
        names.emplace_back(e->d_name);
    }
    closedir(d);
    std::sort(names.begin(), names.end()); // sort.Strings: bytewise
    return names;
}

enum WalkRet { W_CONT, W_SKIPDIR };
using WalkFn = std::function<WalkRet(const std::string &, const struct stat &)>;

WalkRet walk_rec(const std::string &path, const struct stat &st, const WalkFn &fn)
{
    if (!S_ISDIR(st.st_mode))
        return fn(path, st);
    std::vector<std::string> names = sorted_names(path);
    WalkRet r = fn(path, st);
    if (r != W_CONT)
        return r;
    for (const auto &n : names) {
        const std::string fnm = path == "/" ? "/" + n : path + "/" + n;
        struct stat cst;
        if (lstat(fnm.c_str(), &cst) != 0)
            throw HostError(errno_str("lstat", fnm));
        r = walk_rec(fnm, cst, fn);
        if (r != W_CONT) {
            if (!S_ISDIR(cst.st_mode) || r != W_SKIPDIR)
                return r;
        }
    }
    return W_CONT;
}

void go_walk(const std::string &root, const WalkFn &fn)
{
    struct stat st;
    if (lstat(root.c_str(), &st) != 0)
        throw HostError(errno_str("lstat", root));
    walk_rec(root, st, fn);
}

bool has_meta(const std::string &p) { return p.find_first_of("*?[\\") != std::string::npos; }

bool go_match(const char *pat, const char *name)
{
    // filepath.Match: '*' any run of non-separators, '?' one non-separator, [class], '\\' escape
    while (*pat) {
        if (*pat == '*') {
            while (*pat == '*')
                ++pat;
            for (const char *t = name;; ++t) {
                if (go_match(pat, t))
                    return true;
                if (!*t || *t == '/')
                    return false;
            }
        }
        if (!*name)
            return false;
        if (*pat == '?') {
            if (*name == '/')
                return false;
            ++pat, ++name;
        } else if (*pat == '[') {
            ++pat;
            bool neg = *pat == '^';
            if (neg)
                ++pat;
            bool ok = false;
            while (*pat && *pat != ']') {
                char lo = *pat;
                if (lo == '\\' && pat[1])
                    lo = *++pat;
                char hi = lo;
                if (pat[1] == '-' && pat[2] && pat[2] != ']') {
                    pat += 2;
                    hi = *pat;
                    if (hi == '\\' && pat[1])
                        hi = *++pat;
                }
                if (lo <= *name && *name <= hi)
                    ok = true;
                ++pat;
            }
            if (*pat == ']')
                ++pat;
            if (ok == neg)
                return false;
            ++name;
        } else {
            if (*pat == '\\' && pat[1])
                ++pat;
            if (*pat != *name)
                return false;
            ++pat, ++name;
        }
    }
    return !*name;
}

std::vector<std::string> go_glob(const std::string &pattern)
{
    std::vector<std::string> out;
    if (!has_meta(pattern)) {
        struct stat st;
        if (lstat(pattern.c_str(), &st) == 0)
            out.push_back(pattern);
        return out;
    }
    size_t slash = pattern.rfind('/');
    std::string dir = slash == std::string::npos ? "." : (slash == 0 ? "/" : pattern.substr(0, slash));
    const std::string file = slash == std::string::npos ? pattern : pattern.substr(slash + 1);
    if (dir == pattern)
        return out;
    std::vector<std::string> dirs = has_meta(dir) ? go_glob(dir) : std::vector<std::string>{dir};
    for (const auto &d : dirs) {
        struct stat st;
        if (stat(d.c_str(), &st) != 0 || !S_ISDIR(st.st_mode))
            continue;
        for (const auto &n : sorted_names(d))
            if (go_match(file.c_str(), n.c_str()))
                out.push_back(d == "/" ? "/" + n : d + "/" + n);
    }
    return out;
}

// lib/utils/utils.go:161-163
bool is_special(const struct stat &st)
{
    return S_ISCHR(st.st_mode) || S_ISBLK(st.st_mode) || S_ISFIFO(st.st_mode) || S_ISSOCK(st.st_mode);
}

std::string read_link(const std::string &p)
{
    std::string buf(4096, '\0');
    ssize_t n = readlink(p.c_str(), &buf[0], buf.size());
    if (n < 0)
        throw HostError(errno_str("read link", p));
    buf.resize((size_t)n);
    return buf;
}

// ---------------------------------------------------------------------------------------------------
// context stream (add_copy_step.go:153-184,194-238)
// ---------------------------------------------------------------------------------------------------
struct Seg {
    char kind; // 'P' path bytes, 'L' link target bytes, 'F' file content
    std::string bytes;
    std::string path;
    uint64_t size = 0;
};

std::vector<std::string> resolve_from_paths(const std::string &ctx_dir, const char *const *paths, size_t n)
{
    std::vector<std::string> sources;
    for (size_t i = 0; i < n; ++i) {
        const std::string src = go_join(ctx_dir, paths[i]);
        std::vector<std::string> m = go_glob(src);
        if (m.empty())
            sources.push_back(src);
        else
            sources.insert(sources.end(), m.begin(), m.end());
    }
    return sources;
}

std::vector<Seg> context_segments(const std::string &ctx_dir, const char *const *paths, size_t n)
{
    std::vector<Seg> segs;
    for (const auto &source : resolve_from_paths(ctx_dir, paths, n)) {
        try {
            go_walk(source, [&](const std::string &path, const struct stat &st) -> WalkRet {
                if (is_special(st))
                    return S_ISDIR(st.st_mode) ? W_SKIPDIR : W_CONT;
                Seg p;
                p.kind = 'P';
                p.bytes = go_rel(ctx_dir, path);
                segs.push_back(std::move(p));
                if (S_ISDIR(st.st_mode))
                    return W_CONT;
                if (S_ISLNK(st.st_mode)) {
                    Seg l;
                    l.kind = 'L';
                    l.bytes = read_link(path);
                    segs.push_back(std::move(l));
                    return W_CONT;
                }
                Seg f;
                f.kind = 'F';
                f.path = path;
                f.size = (uint64_t)st.st_size;
                segs.push_back(std::move(f));
                return W_CONT;
            });
        } catch (const HostError &e) {
            throw HostError("walk " + source + ": " + e.what());
        }
    }
    return segs;
}

// ---------------------------------------------------------------------------------------------------
// parallel file reader: fills arena memory with pread()
// ---------------------------------------------------------------------------------------------------
struct ReadJob {
    std::string path;
    uint64_t file_off;
    uint64_t len;
    uint8_t *dst;
};

void run_reads(const std::vector<ReadJob> &jobs, int n_threads)
{
    if (jobs.empty())
        return;
    if (n_threads <= 0)
        n_threads = (int)std::min<unsigned>(16, std::max(1u, std::thread::hardware_concurrency()));
    n_threads = (int)std::min<size_t>((size_t)n_threads, jobs.size());
    std::atomic<size_t> next{0};
    std::atomic<bool> failed{false};
    std::string first_err;
    std::mutex mu_obj, *mu = &mu_obj;
    auto worker = [&]() {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= jobs.size() || failed.load())
                return;
            const ReadJob &j = jobs[i];
            int fd = open(j.path.c_str(), O_RDONLY | O_CLOEXEC);
            std::string err;
            if (fd < 0) {
                err = errno_str("open", j.path);
            } else {
                uint64_t done = 0;
                while (done < j.len) {
                    ssize_t r = pread(fd, j.dst + done, j.len - done, (off_t)(j.file_off + done));
                    if (r < 0) {
                        err = errno_str("read", j.path);
                        break;
                    }
                    if (r == 0) {
                        err = "copy file " + j.path + " to tar writer: unexpected EOF"; // io.CopyN
                        break;
                    }
                    done += (uint64_t)r;
                }
                close(fd); // the reference leaks this fd (add_copy_step.go:230-237)
            }
            if (!err.empty()) {
                std::lock_guard<std::mutex> g(*mu);
                if (!failed.exchange(true))
                    first_err = err;
                return;
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t)
        th.emplace_back(worker);
    worker();
    for (auto &t : th)
        t.join();

    if (failed.load())
        throw HostError(first_err);
}

uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

void ck(mksnap_t *eng, int rc, const char *what)
{
    if (rc != 0)
        throw HostError(std::string(what) + ": " + mksnap_last_error(eng));
}

// ---------------------------------------------------------------------------------------------------
// tar header (go1.14 archive/tar Writer.WriteHeader via tario.WriteHeader)
// ---------------------------------------------------------------------------------------------------
struct Hdr {
    std::string name, linkname;
    int64_t mode = 0, uid = 0, gid = 0, size = 0;
    __int128 mtime_ns = 0; // time.Time holds any int64 second count: a base-256 mtime field can exceed int64 nanoseconds
    char typeflag = '0';
};

bool is_ascii(const std::string &s)
{
    for (unsigned char c : s)
        if (c >= 0x80)
            return false;
    return true;
}
std::string to_ascii(const std::string &s)
{
    std::string o;
    for (unsigned char c : s)
        if (c < 0x80)
            o.push_back((char)c);
    return o;
}
bool split_ustar(const std::string &name, std::string &prefix, std::string &suffix)
{
    size_t length = name.size();
    if (length <= 100 || !is_ascii(name))
        return false;
    if (length > 156)
        length = 156;
    else if (name[length - 1] == '/')
        --length;
    size_t i = name.substr(0, length).rfind('/');
    if (i == std::string::npos || i == 0)
        return false;
    size_t nlen = name.size() - i - 1, plen = i;
    if (nlen > 100 || nlen == 0 || plen > 155)
        return false;
    prefix = name.substr(0, i);
    suffix = name.substr(i + 1);
    return true;
}
void fmt_string(uint8_t *b, size_t size, const std::string &s)
{
    size_t n = std::min(s.size(), size);
    memcpy(b, s.data(), n);
    if (s.size() < size)
        b[s.size()] = 0;
    if (s.size() > size && b[size - 1] == '/') {
        size_t k = size;
        while (k > 0 && s[k - 1] == '/')
            --k;
        b[k] = 0;
    }
}
bool fits_octal(size_t n, int64_t x) { return x >= 0 && (uint64_t)x < (1ull << ((n - 1) * 3)); }
void fmt_octal(uint8_t *b, size_t size, int64_t x)
{
    if (!fits_octal(size, x))
        x = 0;
    char tmp[32];
    snprintf(tmp, sizeof tmp, "%llo", (unsigned long long)x);
    std::string s = tmp;
    if (size > s.size() + 1)
        s = std::string(size - s.size() - 1, '0') + s;
    fmt_string(b, size, s);
}
void finish_block(uint8_t *b)
{
    memcpy(b + 257, "ustar\0", 6);
    memcpy(b + 263, "00", 2);
    memset(b + 148, ' ', 8);
    uint32_t chk = 0;
    for (int i = 0; i < 512; ++i)
        chk += b[i];
    fmt_octal(b + 148, 7, chk);
    b[155] = ' ';
}
void template_v7plus(uint8_t *b, const Hdr &h, const std::string &name, const std::string &link, bool ascii_only)
{
    memset(b, 0, 512);
    b[156] = (uint8_t)h.typeflag;
    fmt_string(b + 0, 100, ascii_only ? to_ascii(name) : name);
    fmt_string(b + 157, 100, ascii_only ? to_ascii(link) : link);
    fmt_octal(b + 100, 8, h.mode);
    fmt_octal(b + 108, 8, h.uid);
    fmt_octal(b + 116, 8, h.gid);
    fmt_octal(b + 124, 12, h.size);
    fmt_octal(b + 136, 12, (int64_t)(h.mtime_ns / 1000000000ll));
    fmt_string(b + 265, 32, "");
    fmt_string(b + 297, 32, "");
    fmt_octal(b + 329, 8, 0);
    fmt_octal(b + 337, 8, 0);
}
std::string pax_record(const std::string &k, const std::string &v)
{
    size_t size = k.size() + v.size() + 3;
    size += std::to_string(size).size();
    std::string rec = std::to_string(size) + " " + k + "=" + v + "\n";
    if (rec.size() != size) {
        size = rec.size();
        rec = std::to_string(size) + " " + k + "=" + v + "\n";
    }
    return rec;
}

// returns header bytes (512 or PAX 512 + data + 512)
std::string encode_header(Hdr h)
{
    h.mtime_ns = (h.mtime_ns / 1000000000ll) * 1000000000ll; // write.go:61 Truncate(1s); Writer's Round is then a no-op
    std::map<std::string, std::string> pax;
    bool ustar_ok = true;
    auto verify_string = [&](const std::string &s, size_t size, const char *key) {
        bool too_long = s.size() > size;
        if (!is_ascii(s) || too_long) {
            std::string a, b;
            if (!(key && !strcmp(key, "path") && split_ustar(s, a, b)))
                ustar_ok = false;
            if (!key)
                throw HostError("archive/tar: header field cannot be encoded");
            pax[key] = s;
        }
    };
    auto verify_numeric = [&](int64_t n, size_t size, const char *key) {
        if (!fits_octal(size, n)) {
            ustar_ok = false;
            if (!key)
                throw HostError("archive/tar: header field too long");
            pax[key] = std::to_string(n);
        }
    };
    verify_string(h.name, 100, "path");
    verify_string(h.linkname, 100, "linkpath");
    verify_numeric(h.mode, 8, nullptr);
    verify_numeric(h.uid, 8, "uid");
    verify_numeric(h.gid, 8, "gid");
    verify_numeric(h.size, 12, "size");
    verify_numeric((int64_t)(h.mtime_ns / 1000000000ll), 12, "mtime");
    if ((h.typeflag == '0' || h.typeflag == '3' || h.typeflag == '4' || h.typeflag == '6') && !h.name.empty() &&
        h.name.back() == '/')
        throw HostError("archive/tar: filename may not have trailing slash");
    uint8_t blk[512];
    if (ustar_ok) {
        std::string prefix, name = h.name, sfx;
        if (split_ustar(h.name, prefix, sfx))
            name = sfx;
        else
            prefix.clear();
        template_v7plus(blk, h, name, h.linkname, false);
        fmt_string(blk + 345, 155, prefix);
        finish_block(blk);
        return std::string((const char *)blk, 512);
    }
    std::string out;
    if (!pax.empty()) {
        std::string data;
        for (const auto &kv : pax) // std::map iterates in sorted key order
            data += pax_record(kv.first, kv.second);
        size_t sl = h.name.rfind('/');
        std::string dir = sl == std::string::npos ? "" : h.name.substr(0, sl + 1);
        std::string file = sl == std::string::npos ? h.name : h.name.substr(sl + 1);
        std::string xname = to_ascii(go_join(go_join(dir, "PaxHeaders.0"), file));
        if (xname.size() > 100)
            xname.resize(100);
        while (!xname.empty() && xname.back() == '/')
            xname.pop_back();
        memset(blk, 0, 512);
        blk[156] = 'x';
        fmt_string(blk, 100, xname);
        fmt_octal(blk + 100, 8, 0);
        fmt_octal(blk + 108, 8, 0);
        fmt_octal(blk + 116, 8, 0);
        fmt_octal(blk + 124, 12, (int64_t)data.size());
        fmt_octal(blk + 136, 12, 0);
        finish_block(blk);
        out.assign((const char *)blk, 512);
        out += data;
        out.append((512 - data.size() % 512) % 512, '\0');
    }
    template_v7plus(blk, h, h.name, h.linkname, true);
    finish_block(blk);
    out.append((const char *)blk, 512);
    return out;
}

// ---------------------------------------------------------------------------------------------------
// go1.14 archive/tar: Header.FileInfo().Mode() and the Reader (reader.go Next/readHeader/parsePAX/mergePAX,
// strconv.go parseNumeric/parseOctal/parsePAXTime, format.go getFormat).  Used by UpdateFromTarReader.
// Sparse members (GNU 'S', PAX GNU.sparse.*) are rejected: docker layers do not carry them.
// ---------------------------------------------------------------------------------------------------
int64_t floor_sec(__int128 ns) // Time.Truncate(1s) compares equal iff the floored seconds are equal
{
    __int128 s = ns / 1000000000ll;
    if (ns % 1000000000ll < 0)
        --s;
    return (int64_t)s;
}

constexpr uint32_t GO_MODE_DIR = 1u << 31, GO_MODE_SYMLINK = 1u << 27, GO_MODE_DEVICE = 1u << 26,
                   GO_MODE_NAMED_PIPE = 1u << 25, GO_MODE_SOCKET = 1u << 24, GO_MODE_SETUID = 1u << 23,
                   GO_MODE_SETGID = 1u << 22, GO_MODE_CHAR_DEVICE = 1u << 21, GO_MODE_STICKY = 1u << 20;

uint32_t go_file_mode(const Hdr &h)
{
    const uint32_t m = (uint32_t)h.mode; // os.FileMode(h.Mode)
    uint32_t mode = m & 0777;
    if (m & 04000) mode |= GO_MODE_SETUID;
    if (m & 02000) mode |= GO_MODE_SETGID;
    if (m & 01000) mode |= GO_MODE_STICKY;
    switch (m & ~07777u) {
    case 040000: mode |= GO_MODE_DIR; break;
    case 010000: mode |= GO_MODE_NAMED_PIPE; break;
    case 0120000: mode |= GO_MODE_SYMLINK; break;
    case 060000: mode |= GO_MODE_DEVICE; break;
    case 020000: mode |= GO_MODE_DEVICE | GO_MODE_CHAR_DEVICE; break;
    case 0140000: mode |= GO_MODE_SOCKET; break;
    default: break;
    }
    switch (h.typeflag) {
    case '2': mode |= GO_MODE_SYMLINK; break;
    case '3': mode |= GO_MODE_DEVICE | GO_MODE_CHAR_DEVICE; break;
    case '4': mode |= GO_MODE_DEVICE; break;
    case '5': mode |= GO_MODE_DIR; break;
    case '6': mode |= GO_MODE_NAMED_PIPE; break;
    default: break;
    }
    return mode;
}

bool hdr_is_special(const Hdr &h) // utils.IsSpecialFile(hdr.FileInfo())
{
    return (go_file_mode(h) & (GO_MODE_CHAR_DEVICE | GO_MODE_DEVICE | GO_MODE_NAMED_PIPE | GO_MODE_SOCKET)) != 0;
}

struct TarErr : HostError {
    using HostError::HostError;
};
[[noreturn]] void err_header() { throw TarErr("archive/tar: invalid tar header"); }

std::string c_string(const uint8_t *b, size_t n)
{
    size_t k = 0;
    while (k < n && b[k])
        ++k;
    return std::string((const char *)b, k);
}

int64_t parse_octal(const uint8_t *b, size_t n)
{
    while (n && (b[0] == ' ' || b[0] == 0)) { ++b; --n; }
    while (n && (b[n - 1] == ' ' || b[n - 1] == 0)) --n;
    if (!n)
        return 0;
    size_t k = 0;
    while (k < n && b[k])
        ++k;
    if (!k)
        err_header();
    uint64_t x = 0;
    for (size_t i = 0; i < k; ++i) {
        if (b[i] < '0' || b[i] > '7' || (x >> 61))
            err_header();
        x = x * 8 + (b[i] - '0');
    }
    return (int64_t)x;
}

int64_t parse_numeric(const uint8_t *b, size_t n)
{
    if (n && (b[0] & 0x80)) { // base-256, two's complement when 0x40 is set
        const uint8_t inv = (b[0] & 0x40) ? 0xFF : 0x00;
        uint64_t x = 0;
        for (size_t i = 0; i < n; ++i) {
            uint8_t c = b[i] ^ inv;
            if (i == 0)
                c &= 0x7F;
            if (x >> 56)
                err_header();
            x = (x << 8) | c;
        }
        if (x >> 63)
            err_header();
        return inv ? ~(int64_t)x : (int64_t)x;
    }
    return parse_octal(b, n);
}

int64_t parse_int10(const std::string &v)
{
    if (v.empty())
        err_header();
    size_t i = 0;
    bool neg = false;
    if (v[0] == '-' || v[0] == '+') { neg = v[0] == '-'; i = 1; }
    if (i == v.size())
        err_header();
    int64_t x = 0;
    for (; i < v.size(); ++i) {
        if (v[i] < '0' || v[i] > '9' || x > (INT64_MAX - 9) / 10)
            err_header();
        x = x * 10 + (v[i] - '0');
    }
    return neg ? -x : x;
}

__int128 parse_pax_time(const std::string &s) // seconds[.fraction] -> ns, fraction truncated to 9 digits
{
    const size_t dot = s.find('.');
    const std::string ss = s.substr(0, dot), sn = dot == std::string::npos ? "" : s.substr(dot + 1);
    const __int128 secs = parse_int10(ss);
    if (sn.empty())
        return secs * 1000000000ll;
    int64_t ns = 0;
    for (size_t i = 0; i < 9; ++i) {
        char c = i < sn.size() ? sn[i] : '0';
        if (c < '0' || c > '9')
            err_header();
        ns = ns * 10 + (c - '0');
    }
    for (size_t i = 9; i < sn.size(); ++i)
        if (sn[i] < '0' || sn[i] > '9')
            err_header();
    return (!ss.empty() && ss[0] == '-') ? secs * 1000000000ll - ns : secs * 1000000000ll + ns;
}

std::map<std::string, std::string> parse_pax_records(const uint8_t *b, size_t n)
{
    std::map<std::string, std::string> out;
    size_t pos = 0;
    while (pos < n) {
        size_t sp = pos;
        uint64_t len = 0;
        while (sp < n && b[sp] != ' ') {
            if (b[sp] < '0' || b[sp] > '9' || len > (1ull << 40))
                err_header();
            len = len * 10 + (b[sp] - '0');
            ++sp;
        }
        if (sp == n || sp == pos || len < 5 || len > n - pos)
            err_header();
        const uint8_t *rec = b + sp + 1;
        const size_t rlen = pos + len - (sp + 1);
        if (rlen == 0 || rec[rlen - 1] != '\n')
            err_header();
        const uint8_t *eq = (const uint8_t *)memchr(rec, '=', rlen - 1);
        if (!eq)
            err_header();
        std::string key((const char *)rec, eq - rec), val((const char *)eq + 1, rec + rlen - 1 - (eq + 1));
        if (key.compare(0, 11, "GNU.sparse.") == 0)
            throw TarErr("archive/tar: sparse entries are not supported");
        if (!val.empty())
            out[key] = val;
        else
            out.erase(key);
        pos += len;
    }
    return out;
}

struct TarMember {
    Hdr hdr;
    uint64_t data_off = 0; // offset of the data in the tar stream
    uint64_t data_len = 0; // bytes the Reader exposes (0 for header-only types)
    uint64_t arena_off = 0; // where the data sits in the arena it was placed in (ingest only)
};

// where the stream comes from / goes to: memory (describe) or fd -> pinned arenas (ingest)
struct TarSource {
    virtual ~TarSource() = default;
    virtual bool read_header(uint8_t out[512]) = 0;                                 // false: clean EOF at a block boundary
    virtual const uint8_t *place(const uint8_t hdr[512], uint64_t nb, bool file_content, uint64_t *arena_off) = 0; // header + padded body, contiguous
    virtual void end_marker(const uint8_t zero[512]) = 0;                          // first zero block seen
    virtual uint64_t stream_pos() const = 0;                                        // bytes consumed so far
};

bool is_header_only(char t) { return t == '1' || t == '2' || t == '3' || t == '4' || t == '5' || t == '6'; }

using OnTarMember = std::function<void(const TarMember &, const uint8_t *body)>; // body valid only during the call

std::vector<TarMember> read_tar(TarSource &src, const OnTarMember &on_member = nullptr)
{
    std::vector<TarMember> members;
    std::map<std::string, std::string> pax;
    std::string gnu_name, gnu_link;
    uint8_t blk[512];
    static const uint8_t zero[512] = {0};
    for (;;) {
        if (!src.read_header(blk))
            return members;
        if (memcmp(blk, zero, 512) == 0) {
            src.end_marker(blk);
            return members;
        }
        const int64_t want = parse_octal(blk + 148, 8);
        int64_t us = 0, sg = 0;
        for (int i = 0; i < 512; ++i) {
            const uint8_t c = (i >= 148 && i < 156) ? (uint8_t)' ' : blk[i];
            us += c;
            sg += (int8_t)c;
        }
        if (want != us && want != sg)
            err_header();
        enum { V7, USTAR, STAR, GNU } fmt = V7;
        if (memcmp(blk + 257, "ustar\0", 6) == 0 && memcmp(blk + 508, "tar\0", 4) == 0)
            fmt = STAR;
        else if (memcmp(blk + 257, "ustar\0", 6) == 0)
            fmt = USTAR;
        else if (memcmp(blk + 257, "ustar ", 6) == 0 && memcmp(blk + 263, " \0", 2) == 0)
            fmt = GNU;
        Hdr h;
        h.name = c_string(blk, 100);
        h.mode = parse_numeric(blk + 100, 8);
        h.uid = parse_numeric(blk + 108, 8);
        h.gid = parse_numeric(blk + 116, 8);
        h.size = parse_numeric(blk + 124, 12);
        h.mtime_ns = (__int128)parse_numeric(blk + 136, 12) * 1000000000ll;
        h.typeflag = (char)blk[156];
        h.linkname = c_string(blk + 157, 100);
        if (fmt != V7) {
            (void)parse_numeric(blk + 329, 8); // devmajor / devminor must parse
            (void)parse_numeric(blk + 337, 8);
            std::string prefix;
            if (fmt == USTAR)
                prefix = c_string(blk + 345, 155);
            else if (fmt == STAR)
                prefix = c_string(blk + 345, 131);
            if (!prefix.empty())
                h.name = prefix + "/" + h.name;
        }
        const bool meta = h.typeflag == 'x' || h.typeflag == 'g' || h.typeflag == 'L' || h.typeflag == 'K';
        if (!meta) { // mergePAX, GNU long names, TypeRegA -- before the final size is known
            if (h.typeflag == 'S')
                throw TarErr("archive/tar: sparse entries are not supported");
            for (const auto &kv : pax) {
                if (kv.first == "path") h.name = kv.second;
                else if (kv.first == "linkpath") h.linkname = kv.second;
                else if (kv.first == "uid") h.uid = parse_int10(kv.second);
                else if (kv.first == "gid") h.gid = parse_int10(kv.second);
                else if (kv.first == "mtime") h.mtime_ns = parse_pax_time(kv.second);
                else if (kv.first == "size") h.size = parse_int10(kv.second);
            }
            if (!gnu_name.empty()) h.name = gnu_name;
            if (!gnu_link.empty()) h.linkname = gnu_link;
            if (h.typeflag == '\0')
                h.typeflag = (!h.name.empty() && h.name.back() == '/') ? '5' : '0';
        }
        const int64_t nb = is_header_only(h.typeflag) ? 0 : h.size;
        if (nb < 0)
            err_header();
        const uint64_t data_off = src.stream_pos();
        uint64_t arena_off = 0;
        const uint8_t *body = src.place(blk, (uint64_t)nb, !meta && h.typeflag == '0' && nb > 0, &arena_off);
        if (meta) {
            if (h.typeflag == 'x') {
                pax = parse_pax_records(body, (size_t)nb);
            } else if (h.typeflag == 'g') {
                (void)parse_pax_records(body, (size_t)nb);
                throw TarErr("unsupported type 1100111"); // Next() returns the global header; IsSimilarHeader rejects it
            } else if (h.typeflag == 'L') {
                gnu_name = c_string(body, (size_t)nb);
            } else {
                gnu_link = c_string(body, (size_t)nb);
            }
            continue;
        }
        TarMember m;
        m.hdr = h;
        m.data_off = data_off;
        m.data_len = (uint64_t)nb;
        m.arena_off = arena_off;
        if (on_member)
            on_member(m, body);
        members.push_back(std::move(m));
        pax.clear();
        gnu_name.clear();
        gnu_link.clear();
    }
}

struct MemTarSource : TarSource {
    const uint8_t *p;
    uint64_t n, pos = 0;
    MemTarSource(const uint8_t *p_, uint64_t n_) : p(p_), n(n_) {}
    bool read_header(uint8_t out[512]) override
    {
        if (pos == n)
            return false;
        if (n - pos < 512)
            throw TarErr("unexpected EOF");
        memcpy(out, p + pos, 512);
        pos += 512;
        return true;
    }
    const uint8_t *place(const uint8_t *, uint64_t nb, bool, uint64_t *arena_off) override
    {
        // the data must be complete; a stream that ends inside the zero padding after it is a clean io.EOF for the
        // go1.14 Reader (next(): tryReadFull of the padding returns io.EOF), i.e. the archive simply ends here
        if (n - pos < nb)
            throw TarErr("unexpected EOF");
        const uint8_t *b = p + pos;
        *arena_off = pos;
        pos = std::min<uint64_t>(n, pos + align_up(nb, 512));
        return b;
    }
    void end_marker(const uint8_t *) override
    {
        static const uint8_t zero[512] = {0};
        if (pos == n)
            return;
        if (n - pos < 512)
            throw TarErr("unexpected EOF");
        if (memcmp(p + pos, zero, 512) != 0)
            err_header();
        pos = n;
    }
    uint64_t stream_pos() const override { return pos; }
};

// ---------------------------------------------------------------------------------------------------
// fileio.Copier (lib/fileio/copy.go:30-400) + CopyOperation.Execute (lib/snapshot/copy_op.go:82-147) +
// evalSymlinks (lib/snapshot/utils.go:249-324): the file copy a COPY/ADD step performs when it modifies the file
// system.  `deferred` (ours): regular-file contents are not copied during the traversal but recorded, so the layer
// packer -- which reads the same sources into its arena anyway -- can write them from memory (one read of the
// context instead of one for the copy and one for the layer).
// ---------------------------------------------------------------------------------------------------
struct CopyOwner {
    bool set = false;
    int64_t uid = 0, gid = 0;
    bool overwrite = false;
};
struct DeferredFile {
    std::string dst;
    struct stat st;
    size_t copier = 0; // which Copier (owner rules) finishes it
};

void ck_sys(int rc, const std::string &what, const std::string &path)
{
    if (rc != 0)
        throw HostError(errno_str(what, path));
}

class Copier
{
  public:
    Copier(std::vector<std::string> blacklist, CopyOwner dir_owner, CopyOwner children_owner,
           std::multimap<std::string, DeferredFile> *deferred = nullptr, size_t tag = 0)
        : blacklist_(std::move(blacklist)), dir_owner_(dir_owner), children_owner_(children_owner), deferred_(deferred), tag_(tag)
    {
    }

    void copy_file(const std::string &source, const std::string &target) // copy.go:122-131
    {
        std::string dir = go_clean(target);
        const size_t sl = dir.rfind('/');
        dir = sl == std::string::npos ? "." : (sl == 0 ? "/" : dir.substr(0, sl));
        mkdir_all(dir);
        copy_one(source, target);
    }

    void copy_dir(const std::string &source, const std::string &target) // copy.go:142-156
    {
        if (blacklisted(source))
            return;
        mkdir_all(target);
        copy_dir_contents(source, target, target);
    }

    // copy.go:195-230 with the bytes coming from memory (or, when `data` is null, from src)
    void finish_regular(const struct stat &st, const std::string &src, const std::string &dst, const uint8_t *data, uint64_t len)
    {
        int rfd = -1;
        if (!data) {
            rfd = open(src.c_str(), O_RDONLY | O_CLOEXEC);
            if (rfd < 0)
                throw HostError(errno_str("open", dst)); // the reference reports dst here (copy.go:198)
        }
        const int wfd = open(dst.c_str(), O_WRONLY | O_CREAT | O_CLOEXEC, 0777);
        if (wfd < 0) {
            if (rfd >= 0)
                close(rfd);
            throw HostError(errno_str("create", dst));
        }
        std::string err;
        if (truncate(dst.c_str(), 0) != 0)
            err = errno_str("truncate", dst);
        auto write_all = [&](const uint8_t *p, uint64_t n) {
            while (n && err.empty()) {
                ssize_t w = write(wfd, p, n);
                if (w < 0) {
                    if (errno == EINTR)
                        continue;
                    err = errno_str("copy " + src + " to", dst);
                    break;
                }
                p += w;
                n -= (uint64_t)w;
            }
        };
        if (err.empty()) {
            if (data) {
                write_all(data, len);
            } else {
                std::vector<uint8_t> buf(1 << 20);
                for (;;) {
                    ssize_t r = read(rfd, buf.data(), buf.size());
                    if (r < 0) {
                        if (errno == EINTR)
                            continue;
                        err = errno_str("copy " + src + " to", dst);
                        break;
                    }
                    if (r == 0)
                        break;
                    write_all(buf.data(), (uint64_t)r);
                    if (!err.empty())
                        break;
                }
            }
        }
        close(wfd);
        if (rfd >= 0)
            close(rfd);
        if (!err.empty())
            throw HostError(err);
        int64_t uid = st.st_uid, gid = st.st_gid;
        if (children_owner_.set && children_owner_.overwrite) {
            uid = children_owner_.uid;
            gid = children_owner_.gid;
        }
        ck_sys(chown(dst.c_str(), (uid_t)uid, (gid_t)gid), "chown", dst);
        ck_sys(chmod(dst.c_str(), st.st_mode & 07777), "chmod", dst); // after chown: setuid/setgid survive
    }

  private:
    std::vector<std::string> blacklist_;
    CopyOwner dir_owner_, children_owner_;
    std::multimap<std::string, DeferredFile> *deferred_;
    size_t tag_;

    bool blacklisted(const std::string &p) const
    {
        const std::string a = abs_path(p);
        const size_t sl = a.rfind('/');
        const std::string dir = (sl == 0 ? std::string("/") : a.substr(0, sl)) + "/";
        for (const auto &anc : blacklist_) {
            const std::string b = abs_path(anc);
            if (a == b || b == "/" || dir.compare(0, b.size() + 1, b + "/") == 0)
                return true;
        }
        return false;
    }

    static bool exists(const std::string &p, struct stat *st = nullptr)
    {
        struct stat tmp;
        if (lstat(p.c_str(), st ? st : &tmp) == 0)
            return true;
        if (errno != ENOENT)
            throw HostError(errno_str("lstat", p));
        return false;
    }

    void copy_one(const std::string &src, const std::string &dst) // copyFile, copy.go:163-193
    {
        struct stat st;
        if (lstat(src.c_str(), &st) != 0)
            throw HostError(errno_str("lstat", src));
        if (blacklisted(src)) {
            // the reference only logs in this branch and carries on
        } else if (is_special(st)) {
            return;
        }
        if (S_ISLNK(st.st_mode)) {
            if (exists(dst))
                ck_sys(remove(dst.c_str()), "remove existing file", dst);
            const std::string target = read_link(src);
            ck_sys(symlink(target.c_str(), dst.c_str()), "write link " + dst + " with content", target);
            return;
        }
        if (exists(dst))
            ck_sys(chmod(dst.c_str(), 0777), "chmod", dst);
        if (deferred_) {
            deferred_->emplace(src, DeferredFile{dst, st, tag_});
            return;
        }
        finish_regular(st, src, dst, nullptr, 0);
    }

    void copy_dir_contents(const std::string &src, const std::string &dst, const std::string &orig_dst) // copy.go:252-283
    {
        for (const auto &name : sorted_names(src)) {
            const std::string cur_src = go_join(src, name);
            if (blacklisted(cur_src) || cur_src == orig_dst)
                continue;
            const std::string cur_dst = go_join(dst, name);
            struct stat st;
            if (lstat(cur_src.c_str(), &st) != 0)
                throw HostError(errno_str("lstat", cur_src));
            if (S_ISDIR(st.st_mode)) {
                copy_dir_one(cur_src, cur_dst);
                copy_dir_contents(cur_src, cur_dst, orig_dst);
            } else {
                copy_one(cur_src, cur_dst);
            }
        }
    }

    void copy_dir_one(const std::string &src, const std::string &dst) // copyDir, copy.go:286-329
    {
        struct stat st, dst_st;
        if (lstat(src.c_str(), &st) != 0)
            throw HostError(errno_str("lstat", src));
        if (!S_ISDIR(st.st_mode))
            throw HostError("source " + src + " is not a directory");
        if (blacklisted(src))
            return;
        if (!exists(dst, &dst_st))
            ck_sys(mkdir(dst.c_str(), st.st_mode & 07777), "mkdir", dst);
        else if (!S_ISDIR(dst_st.st_mode))
            throw HostError("dst is not a directory");
        ck_sys(chmod(dst.c_str(), st.st_mode & 07777), "chmod", dst);
        int64_t uid = st.st_uid, gid = st.st_gid;
        if (children_owner_.set && children_owner_.overwrite) {
            uid = children_owner_.uid;
            gid = children_owner_.gid;
        }
        ck_sys(chown(dst.c_str(), (uid_t)uid, (gid_t)gid), "chown", dst);
    }

    void mkdir_all(const std::string &dst) // copy.go:334-399
    {
        if (dst.empty())
            throw HostError("empty dst directory");
        std::string a = go_clean(dst);
        if (a[0] != '/') {
            char cwd[4096];
            if (!getcwd(cwd, sizeof cwd))
                throw HostError("failed to get absolute path of " + dst);
            a = go_join(cwd, a);
        }
        std::string prev = "/";
        const auto parts = split_path(a);
        for (size_t i = 0; i + 1 < parts.size(); ++i) {
            const std::string cur = go_join(prev, parts[i]);
            if (!exists(cur)) {
                ck_sys(mkdir(cur.c_str(), 0755), "mkdir " + cur + " with default mode 0755", cur);
                ck_sys(chown(cur.c_str(), 0, 0), "chown " + cur + " with default owner (0:0)", cur);
            }
            prev = cur;
        }
        if (!exists(a)) {
            ck_sys(mkdir(a.c_str(), 0755), "mkdir " + a + " with default mode 0755", a);
            if (dir_owner_.set)
                ck_sys(chown(a.c_str(), (uid_t)dir_owner_.uid, (gid_t)dir_owner_.gid), "chown", a);
            else
                ck_sys(chown(a.c_str(), 0, 0), "chown", a);
        } else if (dir_owner_.set && dir_owner_.overwrite) {
            ck_sys(chown(a.c_str(), (uid_t)dir_owner_.uid, (gid_t)dir_owner_.gid), "chown", a);
        }
    }
};

// lib/snapshot/utils.go:249-324
std::string walk_link(const std::string &path, const std::string &root, int &walked, bool &islink)
{
    islink = false;
    if (walked > 255)
        throw HostError("eval symlinks: too many links");
    const std::string full = go_join(root, path);
    struct stat st;
    if (lstat(full.c_str(), &st) != 0)
        throw HostError(errno_str("lstat", full));
    if (!S_ISLNK(st.st_mode))
        return path;
    std::string target = read_link(full);
    const bool has_root = target.compare(0, root.size(), root) == 0;
    if (!has_root && !target.empty() && target[0] == '/')
        throw HostError("link points outside of root: " + full + " -> " + target);
    ++walked;
    islink = true;
    return has_root ? target.substr(root.size()) : target;
}

std::string walk_links(const std::string &path, const std::string &root, int &walked)
{
    const size_t sl = path.rfind('/');
    const std::string dir = sl == std::string::npos ? "" : path.substr(0, sl + 1);
    const std::string file = sl == std::string::npos ? path : path.substr(sl + 1);
    bool islink;
    if (dir.empty())
        return walk_link(file, root, walked, islink);
    if (file.empty()) {
        auto trim = [](std::string s) {
            while (!s.empty() && s.back() == '/')
                s.pop_back();
            return s;
        };
        if (trim(dir) == trim(root))
            return dir;
        return walk_links(dir.substr(0, dir.size() - 1), root, walked);
    }
    const std::string newdir = walk_links(dir, root, walked);
    const std::string np = walk_link(newdir.empty() ? file : go_join(newdir, file), root, walked, islink);
    if (!islink || (!np.empty() && np[0] == '/'))
        return np;
    return go_join(newdir, np);
}

std::string eval_symlinks(std::string p, const std::string &src_root)
{
    if (p.empty())
        return p;
    int walked = 0;
    for (;;) {
        const int before = walked;
        const std::string np = walk_links(p, src_root, walked);
        if (before == walked)
            return abs_path(np);
        p = np;
    }
}

// CopyOperation.Execute for one operation.  op.dst must already be resolved against the working directory.
void execute_copy_op(const mkhost_copy_op &c, uint32_t mode, const std::vector<std::string> &blacklist,
                     std::multimap<std::string, DeferredFile> *deferred, std::vector<Copier> *copiers)
{
    const bool chown_flag = mode & MKHOST_COPY_CHOWN, internal = mode & MKHOST_COPY_INTERNAL,
               preserve = mode & MKHOST_COPY_PRESERVE_OWNER;
    if (chown_flag && preserve)
        throw HostError("both chown and archive are true");
    std::string dst = c.dst ? c.dst : "";
    const bool dir_fmt = (!dst.empty() && dst.back() == '/') || dst == "." || dst == "..";
    if (dst.empty() || dst[0] != '/') {
        if (!c.work_dir || c.work_dir[0] != '/')
            throw HostError("check copy param: dst is not absolute path, must specify absolute working directory");
        const std::string d = go_join(c.work_dir, dst);
        dst = dir_fmt ? d + "/" : d;
    }
    for (size_t k = 0; k < c.n_srcs; ++k) {
        std::string src = eval_symlinks(rel_path(c.srcs[k]), c.src_root);
        src = go_join(c.src_root, src);
        struct stat st;
        if (lstat(src.c_str(), &st) != 0)
            throw HostError(errno_str("lstat", src));
        const std::vector<std::string> bl = internal ? std::vector<std::string>{} : blacklist;
        CopyOwner dir_owner, kids;
        if (chown_flag) {
            dir_owner = CopyOwner{true, c.uid, c.gid, false};
            kids = CopyOwner{true, c.uid, c.gid, true};
        } else if (!internal) {
            dir_owner = CopyOwner{true, 0, 0, false};
            kids = CopyOwner{true, 0, 0, true};
        } else if (preserve) {
            dir_owner = CopyOwner{true, (int64_t)st.st_uid, (int64_t)st.st_gid, false};
        }
        Copier copier(bl, dir_owner, kids, deferred, copiers ? copiers->size() : 0);
        if (S_ISDIR(st.st_mode))
            copier.copy_dir(src, dst);
        else if (dir_fmt)
            copier.copy_file(src, go_join(dst, path_base(src)));
        else
            copier.copy_file(src, dst);
        if (copiers)
            copiers->push_back(copier);
    }
}

// ---------------------------------------------------------------------------------------------------
// MemFS (copy-op path): lib/snapshot/mem_fs.go, mem_layer.go
// ---------------------------------------------------------------------------------------------------
struct MemFile {
    std::string src, dst;
    Hdr hdr;
    bool whiteout = false;  // whiteoutMemFile (mem_layer.go:91-132): header-only entry ".wh.<base>"
    std::string deleted;    // path it deletes (the layer key)
    bool has_digest = false; // SHA-256 of the content the tree believes the file has (MKHOST_FILE_DIGESTS; ours)
    std::array<uint8_t, 32> digest{};
};
struct Suspect { // regular file whose header is "similar" but whose remembered content digest can be checked
    std::string src, dst;
    Hdr hdr;
    std::array<uint8_t, 32> known;
};
struct Node {
    MemFile mf;
    std::map<std::string, std::unique_ptr<Node>> children;
};

class MemFS
{
  public:
    MemFS(const std::string &root, int64_t now_unix, std::vector<std::string> blacklist = {})
        : root_(root), now_(now_unix), blacklist_(std::move(blacklist))
    {
        struct stat st;
        if (lstat(root.c_str(), &st) != 0)
            throw HostError("unable to stat root dir: " + root);
        tree_.mf.src = root;
        tree_.mf.dst = "/";
        tree_.mf.hdr = create_header(root, "/", &st, nullptr);
    }

    // mem_layer.go:152-190
    Hdr create_header(const std::string &src, const std::string &dst, const struct stat *st, const Hdr *from)
    {
        Hdr h;
        if (from) { // tar.FileInfoHeader(hdr.FileInfo())
            h = *from;
            h.mode &= 07777;
            if (h.typeflag != '0')
                h.size = 0;
            h.linkname.clear();
        } else {
            const mode_t m = st->st_mode;
            h.mode = m & 0777;
            if (m & S_ISUID) h.mode |= 04000;
            if (m & S_ISGID) h.mode |= 02000;
            if (m & S_ISVTX) h.mode |= 01000;
            h.mtime_ns = (__int128)st->st_mtim.tv_sec * 1000000000ll + st->st_mtim.tv_nsec;
            h.uid = st->st_uid;
            h.gid = st->st_gid;
            if (S_ISREG(m)) { h.typeflag = '0'; h.size = st->st_size; }
            else if (S_ISDIR(m)) h.typeflag = '5';
            else if (S_ISLNK(m)) h.typeflag = '2';
            else if (S_ISCHR(m)) h.typeflag = '3';
            else if (S_ISBLK(m)) h.typeflag = '4';
            else if (S_ISFIFO(m)) h.typeflag = '6';
            else throw HostError("archive/tar: sockets not supported");
        }
        h.name = rel_path(dst);
        const std::string asrc = abs_path(src);
        if (h.typeflag == '5') {
            if (asrc.empty() || asrc.back() != '/')
                h.name += "/";
        } else if (h.typeflag == '2' && !from) {
            std::string target = read_link(asrc);
            if (!target.empty() && target[0] == '/') {
                if (target.compare(0, root_.size(), root_) != 0)
                    throw HostError("trim symlink root: failed to trim root prefix " + root_ + " from path " + target);
                target = abs_path(target.substr(root_.size()));
            }
            h.linkname = target;
        }
        return h;
    }

    void set_now(int64_t now_unix) { now_ = now_unix; }
    const std::string &root() const { return root_; }
    const std::vector<std::string> &blacklist() const { return blacklist_; }

    std::map<std::string, MemFile> add_layer_by_copy_ops(const mkhost_copy_op *ops, size_t n)
    {
        std::map<std::string, MemFile> layer; // std::map == sort.Strings order (mem_layer.go:232-244)
        for (size_t i = 0; i < n; ++i)
            add_to_layer(layer, ops[i]);
        return layer;
    }

    // AddLayerByScan / createLayerByScan (mem_fs.go:260-270,315-341): metadata diff of the root against the
    // merged tree, whiteouts for children that vanished.  (Mountpoint filtering, utils.go:46-50, is the caller's
    // job here: synthetic roots have none.)
    std::map<std::string, MemFile> add_layer_by_scan(std::vector<Suspect> *suspects = nullptr)
    {
        std::map<std::string, MemFile> layer;
        go_walk(root_, [&](const std::string &src, const struct stat &st) -> WalkRet {
            if (should_skip(src, st) || is_descendant_of_any(src, blacklist_))
                return S_ISDIR(st.st_mode) ? W_SKIPDIR : W_CONT;
            if (src.compare(0, root_.size(), root_) != 0)
                throw HostError("failed to trim root prefix " + root_ + " from path " + src);
            const std::string dst = abs_path(src.substr(root_.size()));
            Hdr hdr = create_header(src, dst, &st, nullptr);
            if (suspects && hdr.typeflag == '0') { // content-aware scan: metadata says "unchanged" -- remember to check
                Node *n = nullptr;
                if (!is_updated(dst, hdr, &n) && n && n->mf.has_digest)
                    suspects->push_back(Suspect{src, dst, hdr, n->mf.digest});
            }
            maybe_add(layer, src, dst, hdr, true);
            return W_CONT;
        });
        return layer;
    }

    // a suspect whose content digest differs from the remembered one joins the layer exactly as if isUpdated had
    // said so (mem_fs.go:440-457)
    void add_changed(std::map<std::string, MemFile> &layer, const Suspect &sp)
    {
        add_ancestors(layer, abs_path(sp.dst), false, 0, 0, 0);
        add_header(layer, sp.src, sp.dst, sp.hdr);
    }

    bool get_digest(const std::string &dst, uint8_t d[32])
    {
        Node *cur = &tree_;
        for (const auto &part : split_path(dst)) {
            auto it = cur->children.find(part);
            if (it == cur->children.end())
                return false;
            cur = it->second.get();
        }
        if (!cur->mf.has_digest)
            return false;
        memcpy(d, cur->mf.digest.data(), 32);
        return true;
    }

    void set_digest(const std::string &dst, const uint8_t d[32])
    {
        Node *cur = &tree_;
        for (const auto &part : split_path(dst)) {
            auto it = cur->children.find(part);
            if (it == cur->children.end())
                return;
            cur = it->second.get();
        }
        cur->mf.has_digest = true;
        memcpy(cur->mf.digest.data(), d, 32);
    }

    // UpdateFromTarReader(r, untar=false) (mem_fs.go:165-255): merge the members of a base-layer tar into the tree,
    // hard links in a second pass.  Nothing is written to disk.
    std::map<std::string, MemFile> update_from_tar(const std::vector<TarMember> &members)
    {
        std::map<std::string, MemFile> layer;
        std::map<std::string, Hdr> hardlinks; // the reference ranges over a Go map: order unspecified, result unaffected
        for (const auto &m : members) {
            Hdr hdr = m.hdr;
            const std::string path = go_join(root_, hdr.name);
            if (path_base(path).compare(0, 8, ".wh..wh.") == 0)
                continue;
            if (is_descendant_of_any(path, blacklist_) || hdr_is_special(hdr))
                continue;
            hdr.name = rel_path(hdr.name);
            if (hdr.typeflag == '1') {
                hdr.linkname = abs_path(hdr.linkname);
                hardlinks[path] = hdr;
            } else {
                maybe_add(layer, abs_path(hdr.name), abs_path(hdr.name), hdr, false);
            }
        }
        for (const auto &kv : hardlinks)
            maybe_add(layer, abs_path(kv.second.name), abs_path(kv.second.name), kv.second, false);
        return layer;
    }

    // UpdateFromTarReader(r, untar=true) (mem_fs.go:165-255): the same merge, and every member is also written under the
    // root (untarOneItem, mem_fs.go:574-716; tario.ApplyHeader, lib/tario/apply.go:26-49); the mtimes of the parent
    // directories are restored at the end.  Streaming: members arrive with their bodies (arena or memory).
    struct Untar {
        std::map<std::string, MemFile> layer;
        std::map<std::string, Hdr> hardlinks;
        std::map<std::string, struct timespec> modtimes;
    };
    void untar_member(Untar &u, const TarMember &m, const uint8_t *body)
    {
        Hdr hdr = m.hdr;
        const std::string path = go_join(root_, hdr.name);
        if (path_base(path).compare(0, 8, ".wh..wh.") == 0)
            return;
        if (is_descendant_of_any(path, blacklist_) || hdr_is_special(hdr))
            return;
        const size_t sl = path.rfind('/');
        const std::string parent = sl == 0 ? "/" : path.substr(0, sl);
        if (!u.modtimes.count(parent)) {
            struct stat st;
            if (lstat(parent.c_str(), &st) != 0)
                throw HostError(errno_str("stat parent dir of " + path, parent));
            u.modtimes[parent] = st.st_mtim;
        }
        hdr.name = rel_path(hdr.name);
        if (hdr.typeflag == '1') {
            hdr.linkname = abs_path(hdr.linkname);
            u.hardlinks[path] = hdr;
            return;
        }
        untar_one_item(path, hdr, body, m.data_len);
        maybe_add(u.layer, abs_path(hdr.name), abs_path(hdr.name), hdr, false);
    }
    std::map<std::string, MemFile> untar_finish(Untar &u)
    {
        for (const auto &kv : u.hardlinks) {
            untar_one_item(kv.first, kv.second, nullptr, 0);
            maybe_add(u.layer, abs_path(kv.second.name), abs_path(kv.second.name), kv.second, false);
        }
        for (const auto &kv : u.modtimes) {
            const struct timespec ts[2] = {kv.second, kv.second};
            if (utimensat(AT_FDCWD, kv.first.c_str(), ts, 0) != 0)
                throw HostError(errno_str("chtimes on parent directory", kv.first));
        }
        return std::move(u.layer);
    }

  private:
    std::string root_;
    int64_t now_;
    std::vector<std::string> blacklist_;
    Node tree_;

    static mode_t go_chmod_bits(uint32_t fm) // syscallMode (go1.14 os/file_posix.go)
    {
        mode_t m = fm & 0777;
        if (fm & GO_MODE_SETUID) m |= 04000;
        if (fm & GO_MODE_SETGID) m |= 02000;
        if (fm & GO_MODE_STICKY) m |= 01000;
        return m;
    }

    static void remove_all(const std::string &p) // os.RemoveAll
    {
        struct stat st;
        if (lstat(p.c_str(), &st) != 0) {
            if (errno == ENOENT)
                return;
            throw HostError(errno_str("lstat", p));
        }
        if (S_ISDIR(st.st_mode)) {
            for (const auto &n : sorted_names(p))
                remove_all(go_join(p, n));
            if (rmdir(p.c_str()) != 0)
                throw HostError(errno_str("remove", p));
        } else if (unlink(p.c_str()) != 0) {
            throw HostError(errno_str("remove", p));
        }
    }

    static void apply_header(const std::string &path, const Hdr &hdr) // lib/tario/apply.go:26-49
    {
        struct stat st;
        if (lstat(path.c_str(), &st) != 0)
            throw HostError(errno_str("lstat", path));
        if (S_ISLNK(st.st_mode) || (go_file_mode(hdr) & GO_MODE_SYMLINK))
            throw HostError("update symlink instead of file: " + path);
        ck_sys(chown(path.c_str(), (uid_t)hdr.uid, (gid_t)hdr.gid), "chown", path);
        ck_sys(chmod(path.c_str(), go_chmod_bits(go_file_mode(hdr))), "chmod", path); // after chown: setuid/setgid survive
        __int128 sec = hdr.mtime_ns / 1000000000ll, ns = hdr.mtime_ns % 1000000000ll;
        if (ns < 0) {
            ns += 1000000000ll;
            --sec;
        }
        const struct timespec t = {(time_t)sec, (long)ns};
        const struct timespec ts[2] = {t, t};
        ck_sys(utimensat(AT_FDCWD, path.c_str(), ts, 0), "chtimes", path);
    }

    void untar_one_item(const std::string &path, const Hdr &hdr, const uint8_t *body, uint64_t len) // mem_fs.go:574-650
    {
        const std::string base = path_base(path);
        const size_t sl = path.rfind('/');
        const std::string dir = sl == 0 ? "/" : path.substr(0, sl);
        if (base.compare(0, 4, ".wh.") == 0) { // untarWhiteout
            remove_all(go_join(dir, base.substr(4)));
            return;
        }
        struct stat st;
        if (lstat(path.c_str(), &st) == 0) {
            Hdr local = create_local_header(path, st);
            if (is_similar(local, hdr))
                return; // already on disk
            if ((go_file_mode(hdr) & GO_MODE_DIR) && S_ISDIR(st.st_mode)) {
                apply_header(path, hdr); // existing directories are updated, never deleted
                return;
            }
            remove_all(path);
        } else if (errno != ENOENT) {
            throw HostError(errno_str("lstat", path));
        }
        switch (hdr.typeflag) {
        case '5':
            ck_sys(mkdir(path.c_str(), go_chmod_bits(go_file_mode(hdr))), "create dir", path);
            apply_header(path, hdr);
            break;
        case '2': {
            std::string target = hdr.linkname;
            if (!target.empty() && target[0] == '/')
                target = go_join(root_, target);
            ck_sys(symlink(target.c_str(), path.c_str()), "create symlink " + path + " =>", target);
            ck_sys(lchown(path.c_str(), (uid_t)hdr.uid, (gid_t)hdr.gid), "lchown symlink", path);
            break;
        }
        case '1': {
            const std::string target = go_join(root_, hdr.linkname);
            ck_sys(link(target.c_str(), path.c_str()), "create link " + path + " =>", target);
            apply_header(path, hdr);
            break;
        }
        default: {
            const int fd = open(path.c_str(), O_CREAT | O_TRUNC | O_WRONLY | O_CLOEXEC, go_chmod_bits(go_file_mode(hdr)));
            if (fd < 0)
                throw HostError(errno_str("open file", path));
            uint64_t done = 0;
            while (done < len) {
                ssize_t w = write(fd, body + done, len - done);
                if (w < 0) {
                    if (errno == EINTR)
                        continue;
                    const std::string e = errno_str("read from file", path);
                    close(fd);
                    throw HostError(e);
                }
                done += (uint64_t)w;
            }
            close(fd);
            apply_header(path, hdr);
        }
        }
    }

    // tar.FileInfoHeader(localInfo, linkTarget) as untarOneItem builds it (mem_fs.go:588-605)
    Hdr create_local_header(const std::string &path, const struct stat &st)
    {
        Hdr h;
        const mode_t m = st.st_mode;
        h.mode = m & 0777;
        if (m & S_ISUID) h.mode |= 04000;
        if (m & S_ISGID) h.mode |= 02000;
        if (m & S_ISVTX) h.mode |= 01000;
        h.mtime_ns = (__int128)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec;
        h.uid = st.st_uid;
        h.gid = st.st_gid;
        h.name = path_base(path);
        if (S_ISREG(m)) { h.typeflag = '0'; h.size = st.st_size; }
        else if (S_ISDIR(m)) h.typeflag = '5';
        else if (S_ISLNK(m)) {
            h.typeflag = '2';
            std::string target = read_link(path);
            if (!target.empty() && target[0] == '/') {
                if (target.compare(0, root_.size(), root_) != 0)
                    throw HostError("trim link: failed to trim root prefix " + root_ + " from path " + target);
                target = abs_path(target.substr(root_.size()));
            }
            h.linkname = target;
        } else if (S_ISCHR(m)) h.typeflag = '3';
        else if (S_ISBLK(m)) h.typeflag = '4';
        else if (S_ISFIFO(m)) h.typeflag = '6';
        else throw HostError("archive/tar: sockets not supported");
        return h;
    }

    // pathutils.IsDescendantOfAny (lib/pathutils/path.go:24-36)
    static bool is_descendant_of_any(const std::string &path, const std::vector<std::string> &ancestors)
    {
        const std::string p = abs_path(path);
        size_t sl = p.rfind('/');
        const std::string dir = (sl == 0 ? std::string("/") : p.substr(0, sl)) + "/";
        for (const auto &anc : ancestors) {
            const std::string a = abs_path(anc);
            if (p == a || a == "/" || dir.compare(0, a.size() + 1, a + "/") == 0)
                return true;
        }
        return false;
    }

    void tree_delete(const std::string &path)
    {
        Node *node = &tree_;
        auto parts = split_path(path);
        for (size_t i = 0; i < parts.size(); ++i) {
            auto it = node->children.find(parts[i]);
            if (it != node->children.end()) {
                if (i + 1 == parts.size())
                    node->children.erase(it);
                else
                    node = it->second.get();
            } else if (i + 1 != parts.size()) {
                throw HostError("missing intermediate dir " + parts[i] + " in " + path);
            }
        }
    }

    static MemFile make_whiteout(const std::string &deleted, const std::string &wpath)
    {
        MemFile mf;
        mf.dst = wpath;
        mf.whiteout = true;
        mf.deleted = deleted;
        mf.hdr.name = rel_path(wpath); // &tar.Header{Name: RelPath(whiteoutPath)}: everything else zero
        mf.hdr.typeflag = '0';         // TypeRegA is promoted to TypeReg by tar.Writer
        return mf;
    }

    // utils.go:37-52 (no blacklist on the copy path; mountpoints are the caller's concern)
    static bool should_skip(const std::string &p, const struct stat &st)
    {
        return path_base(p).compare(0, 8, ".wh..wh.") == 0 || is_special(st);
    }

    void tree_put(const MemFile &mf)
    {
        Node *node = &tree_;
        auto parts = split_path(mf.dst);
        for (size_t i = 0; i < parts.size(); ++i) {
            const bool last = i + 1 == parts.size();
            auto it = node->children.find(parts[i]);
            if (it != node->children.end()) {
                if (last) {
                    auto nn = std::make_unique<Node>();
                    nn->mf = mf;
                    if (mf.hdr.typeflag == '5')
                        nn->children = std::move(it->second->children);
                    it->second = std::move(nn);
                } else {
                    node = it->second.get();
                }
            } else if (last) {
                auto nn = std::make_unique<Node>();
                nn->mf = mf;
                node->children[parts[i]] = std::move(nn);
            } else {
                throw HostError("missing intermediate directory " + parts[i] + " in " + mf.dst);
            }
        }
    }

    void add_header(std::map<std::string, MemFile> &layer, const std::string &src, const std::string &dst, const Hdr &hdr)
    {
        const std::string adst = abs_path(dst);
        const std::string base = path_base(adst);
        if (base.compare(0, 4, ".wh.") == 0) { // mem_layer.go:198-206: a whiteout file found on disk / in a layer
            size_t sl = adst.rfind('/');
            const std::string deleted = abs_path(adst.substr(0, sl + 1) + base.substr(4));
            layer[deleted] = make_whiteout(deleted, adst);
            tree_delete(deleted);
            return;
        }
        MemFile mf;
        mf.src = abs_path(src);
        mf.dst = adst;
        mf.hdr = hdr;
        layer[mf.dst] = mf;
        tree_put(mf);
    }

    // tario/compare.go:24-120
    static bool is_similar(const Hdr &h, const Hdr &nh)
    {
        if (h.name.empty() && nh.name.empty())
            return true;
        const bool teq = floor_sec(h.mtime_ns) == floor_sec(nh.mtime_ns);
        const bool meq = go_file_mode(h) == go_file_mode(nh);
        switch (h.typeflag) {
        case '2': return nh.typeflag == '2' && h.linkname == nh.linkname;
        case '1': return nh.typeflag == '1' && teq && h.linkname == nh.linkname && h.uid == nh.uid && h.gid == nh.gid && meq;
        case '5': return nh.typeflag == '5' && teq && h.uid == nh.uid && h.gid == nh.gid && meq;
        case '0': return nh.typeflag == '0' && teq && h.uid == nh.uid && h.gid == nh.gid && h.size == nh.size && meq;
        default: throw HostError(std::string("unsupported type ") + h.typeflag);
        }
    }

    bool is_updated(const std::string &p, const Hdr &hdr, Node **found)
    {
        *found = nullptr;
        Node *cur = &tree_;
        for (const auto &part : split_path(p)) {
            auto it = cur->children.find(part);
            if (it == cur->children.end())
                return true;
            cur = it->second.get();
        }
        *found = cur;
        return !is_similar(cur->mf.hdr, hdr);
    }

    // mem_fs.go:509-569
    std::string add_ancestors(std::map<std::string, MemFile> &layer, const std::string &dst, bool inclusive, int depth,
                              int64_t uid, int64_t gid)
    {
        if (depth >= 1024)
            throw HostError("symlink loop at " + dst);
        Node *last_ancestor = &tree_, *cur = &tree_;
        auto parts = split_path(dst);
        const size_t end = inclusive ? parts.size() : (parts.empty() ? 0 : parts.size() - 1);
        size_t i = 0;
        for (; i < end; ++i) {
            auto it = cur->children.find(parts[i]);
            if (it == cur->children.end())
                break;
            const MemFile mf = it->second->mf;
            add_header(layer, mf.src, mf.dst, mf.hdr);
            Node *n = cur->children[parts[i]].get();
            if (n->mf.hdr.typeflag == '5') {
                last_ancestor = n;
                cur = n;
            } else if (n->mf.hdr.typeflag == '2') {
                std::string remaining;
                for (size_t k = i + 1; k < parts.size(); ++k)
                    remaining = go_join(remaining, parts[k]);
                return add_ancestors(layer, go_join(n->mf.hdr.linkname, remaining), inclusive, depth + 1, uid, gid);
            }
        }
        for (size_t j = i; j < end; ++j) {
            std::string cp;
            for (size_t k = 0; k <= j; ++k)
                cp = go_join(cp, parts[k]);
            cp = abs_path(cp);
            Hdr hdr = create_header("", cp, nullptr, &last_ancestor->mf.hdr);
            hdr.mtime_ns = (__int128)now_ * 1000000000ll; // clk.Now()
            hdr.uid = uid;
            hdr.gid = gid;
            add_header(layer, "", cp, hdr);
        }
        return dst;
    }

    // mem_fs.go:440-482
    void maybe_add(std::map<std::string, MemFile> &layer, const std::string &src, const std::string &dst, const Hdr &hdr,
                   bool create_whiteout)
    {
        Node *n = nullptr;
        const bool updated = is_updated(dst, hdr, &n);
        // the reference ranges over the children of the node found BEFORE the update replaces it
        std::vector<std::pair<std::string, std::string>> kids; // (dst, src)
        if (create_whiteout && hdr.typeflag == '5' && n)
            for (const auto &kv : n->children)
                kids.emplace_back(kv.second->mf.dst, kv.second->mf.src);
        if (updated && dst != "/") {
            add_ancestors(layer, abs_path(dst), false, 0, 0, 0);
            add_header(layer, src, dst, hdr);
        }
        for (const auto &k : kids) {
            struct stat st;
            if (lstat(k.second.c_str(), &st) == 0)
                continue; // still on disk
            if (errno != ENOENT)
                throw HostError(errno_str("check on disk", k.first));
            const std::string ap = abs_path(k.first);
            const std::string base = path_base(ap);
            if (base.compare(0, 4, ".wh.") == 0)
                throw HostError("base name contains whiteout prefix: " + k.first);
            size_t sl = ap.rfind('/');
            const std::string wpath = go_join(ap.substr(0, sl + 1), ".wh." + base);
            layer[k.first] = make_whiteout(k.first, wpath);
            tree_delete(k.first);
            add_ancestors(layer, k.first, false, 0, 0, 0);
        }
    }

    // mem_fs.go:353-420
    void add_to_layer(std::map<std::string, MemFile> &layer, const mkhost_copy_op &c)
    {
        if (c.n_srcs == 0)
            throw HostError("check copy param: srcs cannot be empty");
        std::string dst = c.dst;
        const bool dir_fmt = (!dst.empty() && dst.back() == '/') || dst == "." || dst == "..";
        if (c.n_srcs > 1 && !dir_fmt)
            throw HostError("check copy param: tarring multiple sources, destination must end with \"/\"");
        if (dst.empty() || dst[0] != '/') {
            if (!c.work_dir || c.work_dir[0] != '/')
                throw HostError("check copy param: dst is not absolute path, must specify absolute working directory");
            std::string d = go_join(c.work_dir, dst);
            dst = dir_fmt ? d + "/" : d;
        }
        bool create_dst = true;
        const std::string src_root = c.src_root;
        if (c.n_srcs == 1) {
            const std::string s = go_join(src_root, rel_path(c.srcs[0]));
            struct stat st;
            if (stat(s.c_str(), &st) != 0)
                throw HostError(errno_str("stat src", s));
            if (!S_ISDIR(st.st_mode))
                create_dst = false;
        }
        if (create_dst) {
            std::string resolved = add_ancestors(layer, abs_path(dst), true, 0, c.uid, c.gid);
            if (resolved.empty() || resolved.back() != '/')
                resolved += "/";
            dst = resolved;
        }
        for (size_t k = 0; k < c.n_srcs; ++k) {
            // evalSymlinks (utils.go:249-324): sources inside a build context are not symlinked dirs here
            const std::string src = go_join(src_root, rel_path(c.srcs[k]));
            go_walk(src, [&](const std::string &cur, const struct stat &st) -> WalkRet {
                if (should_skip(cur, st))
                    return S_ISDIR(st.st_mode) ? W_SKIPDIR : W_CONT;
                std::string cur_dst;
                if (cur == src) {
                    if (S_ISDIR(st.st_mode))
                        return W_CONT;
                    cur_dst = dst.back() != '/' ? dst : go_join(dst, path_base(src));
                } else {
                    cur_dst = go_join(dst, cur.substr(src.size()));
                }
                Hdr hdr = create_header(cur, cur_dst, &st, nullptr);
                hdr.uid = c.uid;
                hdr.gid = c.gid;
                maybe_add(layer, cur, cur_dst, hdr, false);
                return W_CONT;
            });
        }
    }
};

void set_err(char *err, size_t n, const std::string &s)
{
    if (err && n) {
        snprintf(err, n, "%s", s.c_str());
    }
}

std::string describe_layer_text(const std::map<std::string, MemFile> &layer)
{
    std::string s;
    char tmp[128];
    for (const auto &kv : layer) {
        const Hdr &h = kv.second.hdr;
        snprintf(tmp, sizeof tmp, "%c %llo %lld %lld %lld %lld ", h.typeflag, (unsigned long long)h.mode, (long long)h.uid,
                 (long long)h.gid, (long long)h.size, (long long)(h.mtime_ns / 1000000000ll));
        s += tmp + kv.second.dst + " " + h.name + " " + kv.second.src + "\n";
    }
    return s;
}

// MemFS.commitLayer (mem_fs.go:424-433) + tario.WriteEntry, with the arena as the tar.Writer sink
void commit_layer(mksnap_t *eng, const std::map<std::string, MemFile> &layer, int n_threads, int tar_fd, uint32_t flags,
                  mkhost_layer_result *out, MemFS *remember = nullptr,
                  const std::function<void(const std::vector<ReadJob> &)> &on_read = nullptr)
{
    const bool want_tar_digest = !(flags & MKHOST_NO_TAR_DIGEST);
    const bool file_digests = remember && (flags & (MKHOST_FILE_DIGESTS | MKHOST_SCAN_CONTENT));
    std::vector<std::string> digest_dst; // stream slot 1+k <-> dst
    std::vector<mksnap_range> rngs;
    ck(eng, mksnap_begin(eng), "begin");
    void *hp = nullptr;
    uint64_t cap = 0;
    int32_t aid = -1;
    uint8_t *a = nullptr;
    uint64_t pos = 0, tar_bytes = 0;
    std::vector<mksnap_extent> ext;
    std::vector<ReadJob> jobs;
    auto acquire = [&]() {
        ck(eng, mksnap_arena_acquire(eng, &hp, &cap, &aid), "arena acquire");
        a = (uint8_t *)hp;
        pos = 0;
        ext.clear();
        jobs.clear();
    };
    // The arena is the tar stream.  A layer larger than one arena goes out in pieces (multiples of 512
    // bytes, so of 64): stream 0 continues across submits, the device keeps the SHA-256 midstate.
    auto flush = [&](bool last) {
        run_reads(jobs, n_threads);
        if (on_read)
            on_read(jobs); // MKHOST_MATERIALIZE: the bytes just read are also the copy's payload
        if (tar_fd >= 0) { // hand the tar bytes on before the arena is recycled
            uint64_t w = 0;
            while (w < pos) {
                ssize_t r = write(tar_fd, a + w, pos - w);
                if (r < 0) {
                    if (errno == EINTR)
                        continue;
                    throw HostError(std::string("write layer tar: ") + strerror(errno));
                }
                w += (uint64_t)r;
            }
        }
        if (want_tar_digest)
            rngs.push_back(mksnap_range{0, pos, 0, last ? 0u : MKSNAP_R_MORE});
        ck(eng, mksnap_arena_submit(eng, aid, pos, ext.data(), ext.size(), rngs.data(), rngs.size()), "arena submit");
        rngs.clear();
        tar_bytes += pos;
        if (!last)
            acquire();
    };
    acquire();
    for (const auto &kv : layer) { // alphabetical order of the layer keys (mem_layer.go:232-244)
        const MemFile &mf = kv.second;
        const std::string hb = encode_header(mf.hdr);
        const uint64_t body = (!mf.whiteout && mf.hdr.typeflag == '0') ? (uint64_t)mf.hdr.size : 0;
        const uint64_t need = hb.size() + align_up(body, 512);
        if (need > cap)
            throw HostError("write diffs: entry " + mf.dst + " (" + std::to_string(need) +
                            " bytes) exceeds the arena; a file is chunked within one arena");
        if (pos + need > cap)
            flush(false);
        memcpy(a + pos, hb.data(), hb.size());
        pos += hb.size();
        if (body) {
            jobs.push_back(ReadJob{mf.src, 0, body, a + pos});
            ext.push_back(mksnap_extent{pos, body, 0, MKSNAP_X_CDC, 0});
            if (file_digests) { // one serial stream per file: slot 1+k (slot 0 is the tar stream)
                digest_dst.push_back(mf.dst);
                rngs.push_back(mksnap_range{pos, body, (uint32_t)digest_dst.size(), 0});
            }
            const uint64_t padded = align_up(body, 512);
            memset(a + pos + body, 0, padded - body);
            pos += padded;
        }
    }
    if (pos + 1024 > cap)
        flush(false);
    memset(a + pos, 0, 1024); // tar.Writer.Close: two zero blocks
    pos += 1024;
    flush(true);
    mksnap_result res;
    ck(eng, mksnap_finish(eng, &res), "finish");
    memset(out->tar_digest, 0, 32);
    if (file_digests && !digest_dst.empty()) {
        std::vector<uint8_t> d((digest_dst.size() + 1) * 32);
        ck(eng, mksnap_get_stream_digests(eng, d.data(), digest_dst.size() + 1), "stream digests");
        if (want_tar_digest)
            memcpy(out->tar_digest, d.data(), 32);
        for (size_t k = 0; k < digest_dst.size(); ++k)
            remember->set_digest(digest_dst[k], d.data() + 32 * (k + 1));
    } else if (want_tar_digest) {
        ck(eng, mksnap_get_stream_digests(eng, out->tar_digest, 1), "stream digests");
    }
    memcpy(out->root, res.root, 32);
    out->n_entries = layer.size();
    out->tar_bytes = tar_bytes;
    out->n_chunks = res.n_chunks;
    out->n_unique = res.n_unique;
}

// SHA-256 of the current content of every suspect (one serial stream per file, thousands in flight; a file larger
// than an arena continues across submits).
std::vector<std::array<uint8_t, 32>> digest_files(mksnap_t *eng, const std::vector<Suspect> &files, int n_threads)
{
    std::vector<std::array<uint8_t, 32>> out(files.size());
    if (files.empty())
        return out;
    ck(eng, mksnap_begin(eng), "begin");
    void *hp = nullptr;
    uint64_t cap = 0, pos = 0;
    int32_t aid = -1;
    std::vector<mksnap_range> rngs;
    std::vector<ReadJob> jobs;
    auto acquire = [&]() {
        ck(eng, mksnap_arena_acquire(eng, &hp, &cap, &aid), "arena acquire");
        cap = cap / 512 * 512;
        pos = 0;
        rngs.clear();
        jobs.clear();
    };
    auto flush = [&]() {
        run_reads(jobs, n_threads);
        ck(eng, mksnap_arena_submit(eng, aid, pos, nullptr, 0, rngs.data(), rngs.size()), "arena submit");
    };
    acquire();
    for (size_t k = 0; k < files.size(); ++k) {
        const uint64_t size = (uint64_t)files[k].hdr.size;
        uint64_t done = 0;
        do {
            if (cap - pos < 512) {
                flush();
                acquire();
            }
            const uint64_t n = std::min<uint64_t>(size - done, cap - pos); // cap - pos is a multiple of 512 (so of 64)
            const bool more = done + n < size;
            if (n)
                jobs.push_back(ReadJob{files[k].src, done, n, (uint8_t *)hp + pos});
            rngs.push_back(mksnap_range{pos, n, (uint32_t)k, more ? MKSNAP_R_MORE : 0u});
            pos = align_up(pos + n, 512);
            done += n;
            if (more) { // at most one piece per stream per submit
                flush();
                acquire();
            }
        } while (done < size);
    }
    flush();
    mksnap_result res;
    ck(eng, mksnap_finish(eng, &res), "finish");
    std::vector<uint8_t> d(files.size() * 32);
    ck(eng, mksnap_get_stream_digests(eng, d.data(), files.size()), "stream digests");
    for (size_t k = 0; k < files.size(); ++k)
        memcpy(out[k].data(), d.data() + 32 * k, 32);
    return out;
}

// UpdateFromTarReader with the GPU in the loop: the tar stream is read from `fd` straight into pinned arenas (the
// arena IS the stream, so SHA-256 over the pieces is the layer's DiffID), members stay contiguous inside one arena so
// every regular file body is one CDC extent.
struct ArenaTarSource : TarSource {
    mksnap_t *eng;
    int fd;
    bool want_digest, file_digests;
    uint8_t *a = nullptr;
    uint64_t cap = 0, pos = 0, consumed = 0, tar_bytes = 0;
    int32_t aid = -1;
    std::vector<mksnap_extent> ext;
    std::vector<mksnap_range> rngs;
    uint32_t n_file_streams = 0; // regular-file members seen so far: member k hashes into stream slot 1+k

    ArenaTarSource(mksnap_t *e, int f, bool d, bool fd_) : eng(e), fd(f), want_digest(d), file_digests(fd_) { acquire(); }

    void acquire()
    {
        void *hp = nullptr;
        ck(eng, mksnap_arena_acquire(eng, &hp, &cap, &aid), "arena acquire");
        a = (uint8_t *)hp;
        cap = cap / 512 * 512;
        pos = 0;
        ext.clear();
        rngs.clear();
    }
    void flush(bool last)
    {
        if (want_digest)
            rngs.push_back(mksnap_range{0, pos, 0, last ? 0u : MKSNAP_R_MORE});
        ck(eng, mksnap_arena_submit(eng, aid, pos, ext.data(), ext.size(), rngs.data(), rngs.size()), "arena submit");
        tar_bytes += pos;
        if (!last)
            acquire();
    }
    uint64_t read_full(uint8_t *dst, uint64_t n)
    {
        uint64_t done = 0;
        while (done < n) {
            ssize_t r = read(fd, dst + done, n - done);
            if (r < 0) {
                if (errno == EINTR)
                    continue;
                throw HostError(std::string("read header: ") + strerror(errno));
            }
            if (r == 0)
                break;
            done += (uint64_t)r;
        }
        consumed += done;
        return done;
    }
    bool eof = false; // the stream ended inside the padding of the last member: clean end (see MemTarSource::place)
    bool read_header(uint8_t out[512]) override
    {
        if (eof)
            return false;
        const uint64_t r = read_full(out, 512);
        if (r == 0)
            return false;
        if (r < 512)
            throw TarErr("unexpected EOF");
        return true;
    }
    const uint8_t *place(const uint8_t hdr[512], uint64_t nb, bool file_content, uint64_t *arena_off) override
    {
        const uint64_t padded = align_up(nb, 512), need = 512 + padded;
        if (need > cap)
            throw HostError("tar member of " + std::to_string(need) + " bytes exceeds the arena; a file is chunked within one arena");
        if (pos + need > cap)
            flush(false);
        memcpy(a + pos, hdr, 512);
        pos += 512;
        const uint64_t got = read_full(a + pos, padded);
        if (got < nb)
            throw TarErr("unexpected EOF");
        *arena_off = pos;
        if (file_content) {
            ext.push_back(mksnap_extent{pos, nb, 0, MKSNAP_X_CDC, 0});
            if (file_digests)
                rngs.push_back(mksnap_range{pos, nb, ++n_file_streams, 0});
        }
        const uint8_t *body = a + pos;
        pos += got;
        if (got < padded)
            eof = true; // ended inside the padding: the go1.14 Reader reports a clean io.EOF on the next Next()
        return body;
    }
    void raw(const uint8_t *p, uint64_t n)
    {
        if (pos + n > cap)
            flush(false);
        memcpy(a + pos, p, n);
        pos += n;
    }
    void end_marker(const uint8_t zero[512]) override
    {
        raw(zero, 512);
        uint8_t nxt[512];
        const uint64_t r = read_full(nxt, 512);
        if (r == 0)
            return;
        if (r < 512)
            throw TarErr("unexpected EOF");
        if (memcmp(nxt, zero, 512) != 0)
            err_header();
        raw(nxt, 512);
        for (;;) { // whatever follows the end marker (record padding) still belongs to the blob that is digested
            if (pos == cap)
                flush(false);
            const uint64_t got = read_full(a + pos, cap - pos);
            pos += got;
            if (got == 0)
                return;
        }
    }
    uint64_t stream_pos() const override { return consumed; }
};

size_t emit_text(const std::string &s, char *out, size_t cap)
{
    if (s.size() + 1 <= cap)
        memcpy(out, s.c_str(), s.size() + 1);
    return s.size() + 1;
}

} // namespace

// persistent snapshot.MemFS mirror (layers accumulate in the merged tree, like context.BuildContext.MemFS)
struct mkhost_memfs {
    MemFS fs;
    mkhost_memfs(const std::string &root, std::vector<std::string> bl) : fs(root, 0, std::move(bl)) {}
};

// =====================================================================================================
extern "C" {

size_t mkhost_encode_tar_header(const mkhost_tar_header *h, uint8_t *out, size_t cap)
{
    try {
        Hdr x;
        x.name = h->name ? h->name : "";
        while (!x.name.empty() && x.name[0] == '/') // write.go:57 strings.TrimLeft(h.Name, "/")
            x.name.erase(0, 1);
        x.linkname = h->linkname ? h->linkname : "";
        x.mode = h->mode; x.uid = h->uid; x.gid = h->gid; x.size = h->size; x.mtime_ns = h->mtime_ns;
        x.typeflag = h->typeflag;
        std::string b = encode_header(x);
        if (b.size() > cap)
            return 0;
        memcpy(out, b.data(), b.size());
        return b.size();
    } catch (const std::exception &) {
        return 0;
    }
}

size_t mkhost_describe_context_stream(const char *context_dir, const char *const *from_paths, size_t n_paths, char *out,
                                      size_t cap, char *err, size_t errlen)
{
    try {
        std::string s;
        for (const auto &g : context_segments(go_clean(context_dir), from_paths, n_paths)) {
            if (g.kind == 'F')
                s += "F " + std::to_string(g.size) + " " + g.path + "\n";
            else
                s += std::string(1, g.kind) + " " + g.bytes + "\n";
        }
        if (s.size() + 1 <= cap)
            memcpy(out, s.c_str(), s.size() + 1);
        return s.size() + 1;
    } catch (const std::exception &e) {
        set_err(err, errlen, std::string("hash context sources: ") + e.what());
        return 0;
    }
}

size_t mkhost_describe_layer(const char *root_dir, int64_t now_unix, const mkhost_copy_op *ops, size_t n_ops, char *out,
                             size_t cap, char *err, size_t errlen)
{
    try {
        MemFS fs(root_dir, now_unix);
        return emit_text(describe_layer_text(fs.add_layer_by_copy_ops(ops, n_ops)), out, cap);
    } catch (const std::exception &e) {
        set_err(err, errlen, std::string("create layer by copy ops: ") + e.what());
        return 0;
    }
}

mkhost_memfs *mkhost_memfs_new(const char *root_dir, const char *const *blacklist, size_t n_blacklist, char *err,
                               size_t errlen)
{
    try {
        std::vector<std::string> bl;
        for (size_t i = 0; i < n_blacklist; ++i)
            bl.emplace_back(blacklist[i]);
        return new mkhost_memfs(root_dir, std::move(bl));
    } catch (const std::exception &e) {
        set_err(err, errlen, e.what());
        return nullptr;
    }
}

void mkhost_memfs_free(mkhost_memfs *m) { delete m; }

int mkhost_memfs_file_digest(mkhost_memfs *m, const char *dst, uint8_t out[32])
{
    return m && dst && m->fs.get_digest(abs_path(dst), out) ? 0 : 1;
}

size_t mkhost_memfs_describe_copy_ops(mkhost_memfs *m, int64_t now_unix, const mkhost_copy_op *ops, size_t n_ops, char *out,
                                      size_t cap, char *err, size_t errlen)
{
    try {
        m->fs.set_now(now_unix);
        return emit_text(describe_layer_text(m->fs.add_layer_by_copy_ops(ops, n_ops)), out, cap);
    } catch (const std::exception &e) {
        set_err(err, errlen, std::string("create layer by copy ops: ") + e.what());
        return 0;
    }
}

size_t mkhost_memfs_describe_scan(mkhost_memfs *m, int64_t now_unix, char *out, size_t cap, char *err, size_t errlen)
{
    try {
        m->fs.set_now(now_unix);
        return emit_text(describe_layer_text(m->fs.add_layer_by_scan()), out, cap);
    } catch (const std::exception &e) {
        set_err(err, errlen, std::string("create layer by scan: ") + e.what());
        return 0;
    }
}

int mkhost_copy_op_execute(const mkhost_copy_op *op, uint32_t mode, const char *const *blacklist, size_t n_blacklist,
                           char *err, size_t errlen)
{
    try {
        if (!op || op->n_srcs == 0)
            throw HostError("check copy param: srcs cannot be empty");
        std::vector<std::string> bl;
        for (size_t i = 0; i < n_blacklist; ++i)
            bl.emplace_back(blacklist[i]);
        std::multimap<std::string, DeferredFile> deferred;
        std::vector<Copier> copiers;
        execute_copy_op(*op, mode, bl, (mode & MKHOST_COPY_DEFERRED) ? &deferred : nullptr, &copiers);
        for (const auto &kv : deferred) { // stand-in for the arena: one read of each source, written from memory
            std::vector<uint8_t> buf((size_t)kv.second.st.st_size);
            std::vector<ReadJob> job{ReadJob{kv.first, 0, (uint64_t)buf.size(), buf.data()}};
            run_reads(job, 1);
            copiers[kv.second.copier].finish_regular(kv.second.st, kv.first, kv.second.dst, buf.data(), buf.size());
        }
        return 0;
    } catch (const std::exception &e) {
        set_err(err, errlen, e.what());
        return -1;
    }
}

int mkhost_memfs_commit_copy_ops(mkhost_memfs *m, mksnap_t *eng, int64_t now_unix, const mkhost_copy_op *ops, size_t n_ops,
                                 int n_threads, int tar_fd, uint32_t flags, mkhost_layer_result *out, char *err,
                                 size_t errlen)
{
    try {
        m->fs.set_now(now_unix);
        if (!(flags & MKHOST_MATERIALIZE)) {
            commit_layer(eng, m->fs.add_layer_by_copy_ops(ops, n_ops), n_threads, tar_fd, flags, out, &m->fs);
            return 0;
        }
        // the copy onto the file system (CopyOperation.Execute) and the layer share one read of the
        // context.  Directories, symlinks and chmods first; regular files are written from the arena as it fills;
        // what the layer did not need to read (entries the tree already holds) is copied from disk at the end.
        std::multimap<std::string, DeferredFile> deferred;
        std::vector<Copier> copiers;
        const uint32_t mode = (flags & MKHOST_MATERIALIZE_CHOWN) ? MKHOST_COPY_CHOWN : 0u;
        for (size_t i = 0; i < n_ops; ++i) {
            mkhost_copy_op on_disk = ops[i];
            std::string dst = ops[i].dst ? ops[i].dst : "";
            const bool dir_fmt = (!dst.empty() && dst.back() == '/') || dst == "." || dst == "..";
            if (dst.empty() || dst[0] != '/') {
                if (!ops[i].work_dir || ops[i].work_dir[0] != '/')
                    throw HostError("check copy param: dst is not absolute path, must specify absolute working directory");
                dst = go_join(ops[i].work_dir, dst);
            }
            dst = go_join(m->fs.root(), dst) + (dir_fmt ? "/" : ""); // image path -> path under the MemFS root ("/" in a real build)
            on_disk.dst = dst.c_str();
            execute_copy_op(on_disk, mode, m->fs.blacklist(), &deferred, &copiers);
        }
        auto write_from_arena = [&](const std::vector<ReadJob> &jobs) {
            for (const auto &j : jobs) {
                if (j.file_off != 0)
                    continue;
                auto range = deferred.equal_range(j.path);
                for (auto it = range.first; it != range.second; ++it)
                    copiers[it->second.copier].finish_regular(it->second.st, j.path, it->second.dst, j.dst, j.len);
                deferred.erase(range.first, range.second);
            }
        };
        commit_layer(eng, m->fs.add_layer_by_copy_ops(ops, n_ops), n_threads, tar_fd, flags, out, &m->fs, write_from_arena);
        for (const auto &kv : deferred)
            copiers[kv.second.copier].finish_regular(kv.second.st, kv.first, kv.second.dst, nullptr, 0);
        return 0;
    } catch (const std::exception &e) {
        set_err(err, errlen, std::string("failed to generate diff layer: ") + e.what());
        return -1;
    }
}

int mkhost_memfs_commit_scan(mkhost_memfs *m, mksnap_t *eng, int64_t now_unix, int n_threads, int tar_fd, uint32_t flags,
                             mkhost_layer_result *out, char *err, size_t errlen)
{
    try {
        m->fs.set_now(now_unix);
        if (!(flags & MKHOST_SCAN_CONTENT)) {
            commit_layer(eng, m->fs.add_layer_by_scan(), n_threads, tar_fd, flags, out, &m->fs);
            return 0;
        }
        // content-aware scan: the reference trusts mtime+size and therefore sync()s and sleeps
        // a second before every scan (mem_fs.go:291-311); here files the metadata calls unchanged are re-hashed on
        // the GPU and compared with the digest remembered when they were committed.
        std::vector<Suspect> suspects;
        std::map<std::string, MemFile> layer = m->fs.add_layer_by_scan(&suspects);
        const std::vector<std::array<uint8_t, 32>> now = digest_files(eng, suspects, n_threads);
        for (size_t k = 0; k < suspects.size(); ++k)
            if (now[k] != suspects[k].known)
                m->fs.add_changed(layer, suspects[k]);
        commit_layer(eng, layer, n_threads, tar_fd, flags, out, &m->fs);
        return 0;
    } catch (const std::exception &e) {
        set_err(err, errlen, std::string("failed to generate diff layer: ") + e.what());
        return -1;
    }
}

size_t mkhost_memfs_describe_update_from_tar_ex(mkhost_memfs *m, int64_t now_unix, int tar_fd, uint32_t flags, char *out,
                                                size_t cap, char *err, size_t errlen)
{
    try {
        m->fs.set_now(now_unix);
        std::vector<uint8_t> buf;
        uint8_t tmp[65536];
        for (;;) {
            ssize_t r = read(tar_fd, tmp, sizeof tmp);
            if (r < 0) {
                if (errno == EINTR)
                    continue;
                throw HostError(std::string("read header: ") + strerror(errno));
            }
            if (r == 0)
                break;
            buf.insert(buf.end(), tmp, tmp + r);
        }
        MemTarSource src(buf.data(), buf.size());
        if (flags & MKHOST_UNTAR) {
            MemFS::Untar u;
            read_tar(src, [&](const TarMember &mem, const uint8_t *body) { m->fs.untar_member(u, mem, body); });
            return emit_text(describe_layer_text(m->fs.untar_finish(u)), out, cap);
        }
        return emit_text(describe_layer_text(m->fs.update_from_tar(read_tar(src))), out, cap);
    } catch (const std::exception &e) {
        set_err(err, errlen, std::string("update memfs from tar: ") + e.what());
        return 0;
    }
}

size_t mkhost_memfs_describe_update_from_tar(mkhost_memfs *m, int64_t now_unix, int tar_fd, char *out, size_t cap, char *err,
                                             size_t errlen)
{
    return mkhost_memfs_describe_update_from_tar_ex(m, now_unix, tar_fd, 0, out, cap, err, errlen);
}

int mkhost_memfs_update_from_tar(mkhost_memfs *m, mksnap_t *eng, int64_t now_unix, int tar_fd, uint32_t flags,
                                 mkhost_layer_result *out, char *err, size_t errlen)
{
    try {
        m->fs.set_now(now_unix);
        const bool want_digest = !(flags & MKHOST_NO_TAR_DIGEST);
        ck(eng, mksnap_begin(eng), "begin");
        const bool file_digests = (flags & MKHOST_FILE_DIGESTS) != 0;
        ArenaTarSource src(eng, tar_fd, want_digest, file_digests);
        const bool untar = (flags & MKHOST_UNTAR) != 0;
        MemFS::Untar u;
        const std::vector<TarMember> members =
            untar ? read_tar(src, [&](const TarMember &mem, const uint8_t *body) { m->fs.untar_member(u, mem, body); }) // files written from the arena
                  : read_tar(src);
        src.flush(true);
        mksnap_result res;
        ck(eng, mksnap_finish(eng, &res), "finish");
        memset(out->tar_digest, 0, 32);
        std::vector<uint8_t> d((size_t)(src.n_file_streams + 1) * 32);
        if (want_digest || src.n_file_streams)
            ck(eng, mksnap_get_stream_digests(eng, d.data(), src.n_file_streams + 1), "stream digests");
        if (want_digest)
            memcpy(out->tar_digest, d.data(), 32);
        memcpy(out->root, res.root, 32);
        const std::map<std::string, MemFile> layer = untar ? m->fs.untar_finish(u) : m->fs.update_from_tar(members);
        out->n_entries = layer.size();
        if (file_digests) { // remember the content digest of every regular member that made it into the tree
            uint32_t k = 0;
            for (const auto &mem : members) {
                if (mem.hdr.typeflag != '0' || mem.data_len == 0)
                    continue;
                ++k;
                const std::string dst = abs_path(mem.hdr.name);
                auto it = layer.find(dst);
                if (it != layer.end() && !it->second.whiteout && it->second.hdr.typeflag == '0')
                    m->fs.set_digest(dst, d.data() + 32 * k);
            }
        }
        out->tar_bytes = src.tar_bytes;
        out->n_chunks = res.n_chunks;
        out->n_unique = res.n_unique;
        return 0;
    } catch (const std::exception &e) {
        set_err(err, errlen, std::string("update memfs from tar: ") + e.what());
        return -1;
    }
}

// ---- cache.Manager wire format (lib/cache/cache_manager.go:34-35,239-252) ----
size_t mkhost_cache_key(const char *cache_id, int chunk_table, char *out, size_t cap)
{
    return emit_text(std::string("makisu_builder_cache_") + (cache_id ? cache_id : "") + (chunk_table ? "_chunks" : ""), out, cap);
}

size_t mkhost_cache_entry_create(const char *tar_hex, const char *gzip_hex, char *out, size_t cap)
{
    if (!tar_hex) // createEntry(nil)
        return emit_text("MAKISU_CACHE_EMPTY", out, cap);
    return emit_text(std::string(tar_hex) + "," + (gzip_hex ? gzip_hex : ""), out, cap);
}

int mkhost_cache_entry_parse(const char *entry, char *tar_digest, size_t tar_cap, char *gzip_digest, size_t gzip_cap,
                             char *err, size_t errlen)
{
    const std::string e = entry ? entry : "";
    const size_t c = e.find(',');
    if (c == std::string::npos) {
        set_err(err, errlen, "parse redis entry: " + e);
        return -1;
    }
    const std::string t = "sha256:" + e.substr(0, c), g = "sha256:" + e.substr(c + 1); // SplitN(entry, ",", 2)
    if (t.size() + 1 > tar_cap || g.size() + 1 > gzip_cap) {
        set_err(err, errlen, "parse redis entry: output buffer too small");
        return -1;
    }
    memcpy(tar_digest, t.c_str(), t.size() + 1);
    memcpy(gzip_digest, g.c_str(), g.size() + 1);
    return 0;
}

size_t mkhost_cache_chunk_entry_create(const uint8_t root[32], uint64_t n_unique, char *out, size_t cap)
{
    static const char *hx = "0123456789abcdef";
    std::string s;
    for (int i = 0; i < 32; ++i) {
        s += hx[root[i] >> 4];
        s += hx[root[i] & 15];
    }
    return emit_text(s + "," + std::to_string(n_unique), out, cap);
}

int mkhost_cache_chunk_entry_parse(const char *entry, uint8_t root[32], uint64_t *n_unique, char *err, size_t errlen)
{
    const std::string e = entry ? entry : "";
    auto nib = [](char c) { return c >= '0' && c <= '9' ? c - '0' : (c >= 'a' && c <= 'f' ? c - 'a' + 10 : -1); };
    bool ok = e.size() > 65 && e[64] == ',';
    for (int i = 0; ok && i < 64; ++i)
        ok = nib(e[i]) >= 0;
    uint64_t n = 0;
    for (size_t i = 65; ok && i < e.size(); ++i) {
        ok = e[i] >= '0' && e[i] <= '9' && n <= (UINT64_MAX - 9) / 10;
        n = n * 10 + (uint64_t)(e[i] - '0');
    }
    if (!ok) {
        set_err(err, errlen, "parse chunk table entry: " + e);
        return -1;
    }
    for (int i = 0; i < 32; ++i)
        root[i] = (uint8_t)(nib(e[2 * i]) * 16 + nib(e[2 * i + 1]));
    *n_unique = n;
    return 0;
}

int mkhost_context_crc32(mksnap_t *eng, const void *prefix, size_t prefix_len, const char *context_dir,
                         const char *const *from_paths, size_t n_paths, int n_threads, uint32_t *crc_out,
                         uint64_t *stream_len_out, char *err, size_t errlen)
{
    try {
        std::vector<Seg> segs = context_segments(go_clean(context_dir), from_paths, n_paths);
        uint64_t total = prefix_len;
        for (const auto &g : segs)
            total += g.kind == 'F' ? g.size : g.bytes.size();
        ck(eng, mksnap_begin(eng), "begin");

        void *hp = nullptr;
        uint64_t cap = 0;
        int32_t aid = -1;
        uint64_t pos = 0;
        std::vector<mksnap_extent> ext;
        std::vector<ReadJob> jobs;
        auto acquire = [&]() {
            ck(eng, mksnap_arena_acquire(eng, &hp, &cap, &aid), "arena acquire");
            pos = 0;
            ext.clear();
            jobs.clear();
        };
        auto flush = [&]() {
            run_reads(jobs, n_threads);
            ck(eng, mksnap_arena_submit(eng, aid, pos, ext.data(), ext.size(), nullptr, 0), "arena submit");
            aid = -1;
        };
        acquire();
        uint64_t after = total; // stream bytes not yet placed
        auto put_bytes = [&](const void *p, size_t n) {
            if (n == 0)
                return;
            uint64_t o = align_up(pos, 16);
            if (o + n > cap) {
                flush();
                acquire();
                o = 0;
            }
            if (n > cap)
                throw HostError("path string larger than the arena");
            memcpy((uint8_t *)hp + o, p, n);
            after -= n;
            ext.push_back(mksnap_extent{o, n, after, MKSNAP_X_CRC, 0});
            pos = o + n;
        };
        put_bytes(prefix, prefix_len);
        for (const auto &g : segs) {
            if (g.kind != 'F') {
                put_bytes(g.bytes.data(), g.bytes.size());
                continue;
            }
            uint64_t done = 0; // CRC is linear: a file may be split across arenas at any 16-byte boundary
            while (done < g.size) {
                uint64_t o = align_up(pos, 512);
                if (o + 4096 > cap) {
                    flush();
                    acquire();
                    o = 0;
                }
                uint64_t n = std::min<uint64_t>(g.size - done, (cap - o) / 16 * 16);
                jobs.push_back(ReadJob{g.path, done, n, (uint8_t *)hp + o});
                after -= n;
                ext.push_back(mksnap_extent{o, n, after, MKSNAP_X_CRC, 0});
                pos = o + n;
                done += n;
            }
        }
        flush();
        mksnap_result res;
        ck(eng, mksnap_finish(eng, &res), "finish");
        if (res.crc_bytes != total)
            throw HostError("internal: stream length mismatch");
        *crc_out = mksnap_ctx_crc32(&res);
        if (stream_len_out)
            *stream_len_out = total;
        return 0;
    } catch (const std::exception &e) {
        set_err(err, errlen, std::string("hash context sources: ") + e.what());
        return -1;
    }
}

int mkhost_commit_copy_ops(mksnap_t *eng, const char *root_dir, int64_t now_unix, const mkhost_copy_op *ops, size_t n_ops,
                           int n_threads, mkhost_layer_result *out, char *err, size_t errlen)
{
    return mkhost_commit_copy_ops_to_fd(eng, root_dir, now_unix, ops, n_ops, n_threads, -1, out, err, errlen);
}

int mkhost_commit_copy_ops_to_fd(mksnap_t *eng, const char *root_dir, int64_t now_unix, const mkhost_copy_op *ops,
                                 size_t n_ops, int n_threads, int tar_fd, mkhost_layer_result *out, char *err,
                                 size_t errlen)
{
    return mkhost_commit_copy_ops_ex(eng, root_dir, now_unix, ops, n_ops, n_threads, tar_fd, 0, out, err, errlen);
}

int mkhost_commit_copy_ops_ex(mksnap_t *eng, const char *root_dir, int64_t now_unix, const mkhost_copy_op *ops,
                              size_t n_ops, int n_threads, int tar_fd, uint32_t flags, mkhost_layer_result *out,
                              char *err, size_t errlen)
{
    try {
        MemFS fs(root_dir, now_unix);
        commit_layer(eng, fs.add_layer_by_copy_ops(ops, n_ops), n_threads, tar_fd, flags, out);
        return 0;
    } catch (const std::exception &e) {
        set_err(err, errlen, std::string("failed to generate diff layer: ") + e.what());
        return -1;
    }
}

} // extern "C"
