// pa_frontend.cpp -- file-level drop-in for the `distribution` path:
// general.inp (Main:103-357 header switches, Main:2882-3666 keyword loop), molecular.inp
// (Molecular:3444-3700), xyz / lmp trajectory readers (Molecular:4827-5063) and the
// distribution input file (Distribution:328-616), with Fortran list-directed token rules.
// Everything numeric is delegated to the C ABI (pa_job_setup, pa_distribution_histogram_multi,
// pa_transfer_histogram, pa_write_distribution).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <map>
#include <fstream>
#include <sstream>
#include <chrono>
#include <array>
#include <algorithm>
#include <functional>
#include <mutex>
#include <deque>
#include <condition_variable>
#include <thread>
#include <fcntl.h>
#include <unistd.h>
#include <sys/stat.h>
#include "pa_internal.h"

namespace {

// ---- Fortran list-directed tokens: blanks/commas separate, quotes group, '/' ends the record
std::vector<std::string> tokens_of(const std::string &line)
{
    std::vector<std::string> out;
    size_t i = 0, n = line.size();
    while (i < n) {
        while (i < n && (line[i] == ' ' || line[i] == '\t' || line[i] == ',' || line[i] == '\r' || line[i] == '\n')) ++i;
        if (i >= n) break;
        if (line[i] == '/') break;
        if (line[i] == '"' || line[i] == '\'') {
            const char q = line[i++];
            std::string t;
            while (i < n) {
                if (line[i] == q) { if (i + 1 < n && line[i + 1] == q) { t += q; i += 2; continue; } ++i; break; }
                t += line[i++];
            }
            out.push_back(t);
        } else {
            std::string t;
            while (i < n && line[i] != ' ' && line[i] != '\t' && line[i] != ',' && line[i] != '/' && line[i] != '\r' && line[i] != '\n') t += line[i++];
            out.push_back(t);
        }
    }
    return out;
}

bool parse_int(const std::string &s, int &v)
{
    char *e = nullptr;
    const long x = strtol(s.c_str(), &e, 10);
    if (e == s.c_str() || *e) return false;
    v = (int)x;
    return true;
}
// Decimal text -> REAL(4), correctly rounded like strtof (what the list-directed READ of the reference produces,
// Molecular:5042), for the coordinates of a trajectory line.  Fast path: [+-]digits[.digits][(e|E)[+-]digits] with at
// most 15 significant digits and a decimal exponent within +-22.  Then the digits are an integer m < 2^53 and 10^|e| is
// exact in double, so d = m * 10^e (or m / 10^-e) is the correctly rounded double of the decimal value, and rounding d
// to float32 equals rounding the decimal value directly unless d sits exactly on the midpoint between two floats
// (d is the nearest double and the midpoint is a double, so value and d lie on the same side of it otherwise).
// Midpoints, tiny magnitudes (float32 subnormals), long digit strings, large exponents, inf / nan / hex and anything
// else unusual go to strtof.  *end receives the first unparsed character, as with strtof.
float parse_coordinate(const char *p, const char **end)
{
    static const double pow10[23] = { 1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16,
                                      1e17, 1e18, 1e19, 1e20, 1e21, 1e22 };
    const char *q = p;
    bool neg = false;
    if (*q == '+' || *q == '-') { neg = (*q == '-'); ++q; }
    unsigned long long m = 0;
    int ndig = 0, nsig = 0, e10 = 0;
    bool ok = true;
    while (*q >= '0' && *q <= '9') { if (nsig || *q != '0') { if (++nsig > 15) ok = false; else m = m * 10 + (unsigned)(*q - '0'); } ++ndig; ++q; }
    if (*q == '.') {
        ++q;
        while (*q >= '0' && *q <= '9') { if (nsig || *q != '0') { if (++nsig > 15) ok = false; else m = m * 10 + (unsigned)(*q - '0'); } --e10; ++ndig; ++q; }
    }
    if (ndig == 0 || *q == 'x' || *q == 'X') ok = false;      // nothing numeric, or a hexadecimal constant: strtof decides
    if (ok && (*q == 'e' || *q == 'E')) {
        const char *x = q + 1;
        bool eneg = false;
        if (*x == '+' || *x == '-') { eneg = (*x == '-'); ++x; }
        if (*x >= '0' && *x <= '9') {
            int ev = 0;
            while (*x >= '0' && *x <= '9') { if (ev < 10000) ev = ev * 10 + (*x - '0'); ++x; }
            e10 += eneg ? -ev : ev;
            q = x;
        }
    }
    if (ok && e10 >= -22 && e10 <= 22) {
        double d = (double)m;
        d = e10 < 0 ? d / pow10[-e10] : d * pow10[e10];
        unsigned long long bits; memcpy(&bits, &d, 8);
        if ((bits & 0x1FFFFFFFull) != 0x10000000ull && (d == 0.0 || d > 1e-30)) {
            *end = q;
            const float f = (float)d;
            return neg ? -f : f;
        }
    }
    char *e;
    const float f = strtof(p, &e);
    *end = e;
    return f;
}

bool parse_real4(const std::string &s, float &v)      // decimal -> REAL(4) directly
{
    std::string t = s;
    for (char &c : t) if (c == 'd' || c == 'D') c = 'e';
    char *e = nullptr;
    v = strtof(t.c_str(), &e);
    return e != t.c_str() && !*e;
}
bool parse_real8(const std::string &s, double &v)
{
    std::string t = s;
    for (char &c : t) if (c == 'd' || c == 'D') c = 'e';
    char *e = nullptr;
    v = strtod(t.c_str(), &e);
    return e != t.c_str() && !*e;
}
bool parse_logical(const std::string &s, bool &v)
{
    size_t i = 0;
    if (i < s.size() && s[i] == '.') ++i;
    if (i >= s.size()) return false;
    if (s[i] == 'T' || s[i] == 't') { v = true; return true; }
    if (s[i] == 'F' || s[i] == 'f') { v = false; return true; }
    return false;
}

bool read_lines(const std::string &path, std::vector<std::string> &lines)
{
    std::ifstream f(path);
    if (!f) return false;
    std::string l;
    while (std::getline(f, l)) lines.push_back(l);
    return true;
}

// ---- element table, Molecular:9-92 (REAL(4) values)
const std::map<std::string, float> &default_masses()
{
    static const std::map<std::string, float> m = {
        {"H", 1.008f}, {"He", 4.002f}, {"Li", 6.94f}, {"Be", 9.012f}, {"B", 10.81f}, {"C", 12.011f}, {"N", 14.007f},
        {"O", 15.999f}, {"F", 18.998f}, {"Ne", 20.1797f}, {"Na", 22.989f}, {"Mg", 24.305f}, {"Al", 26.981f},
        {"Si", 28.085f}, {"P", 30.973762f}, {"S", 32.06f}, {"Cl", 35.45f}, {"Ar", 39.95f}, {"K", 39.0983f},
        {"Ca", 40.078f}, {"Sc", 44.955f}, {"Ti", 47.867f}, {"V", 50.9415f}, {"Cr", 51.9961f}, {"Mn", 54.938f},
        {"Fe", 55.845f}, {"Co", 58.933f}, {"Ni", 58.6934f}, {"Cu", 63.546f}, {"Zn", 65.38f}, {"Ga", 69.723f},
        {"Ge", 72.63f}, {"As", 74.921f}, {"Se", 78.971f}, {"Br", 79.904f}, {"Kr", 83.798f}, {"Rb", 85.4678f},
        {"Sr", 87.62f}, {"Y", 88.905f}, {"Zr", 91.222f}, {"Nb", 92.906f}, {"Mo", 95.95f}, {"Ru", 101.07f},
        {"Rh", 102.905f}, {"Pd", 106.42f}, {"Ag", 107.8682f}, {"Cd", 112.414f}, {"In", 114.818f}, {"Sn", 118.71f},
        {"Sb", 121.76f}, {"Te", 127.6f}, {"I", 126.904f}, {"Xe", 131.293f}, {"Cs", 132.905f}, {"Ba", 137.327f},
        {"La", 138.905f}, {"Ce", 140.116f}, {"Pr", 140.907f}, {"Nd", 144.242f}, {"Sm", 150.36f}, {"Eu", 151.964f},
        {"Gd", 157.249f}, {"Tb", 158.925f}, {"Dy", 162.5f}, {"Ho", 164.93f}, {"Er", 167.259f}, {"Tm", 168.934f},
        {"Yb", 173.045f}, {"Lu", 174.966f}, {"Hf", 178.486f}, {"Ta", 180.947f}, {"W", 183.84f}, {"Re", 186.207f},
        {"Os", 190.23f}, {"Ir", 192.217f}, {"Pt", 195.084f}, {"Au", 196.966f}, {"Hg", 200.592f}, {"Tl", 204.38f},
        {"Pb", 207.2f}, {"Bi", 208.98f}, {"Th", 232.0377f}, {"Pa", 231.035f}, {"U", 238.028f}, {"X", 0.0f}, {"D", 0.0f},
    };
    return m;
}

struct HostType {
    int charge = 0, natoms = 0, nmol = 0;
    float realcharge = 0.f;
    double mass = 0.0;
    std::vector<std::string> elements;
    std::vector<double> atom_mass, atom_charge;
    std::vector<char> mass_manual, charge_manual;
    std::vector<float> xyz;
};

struct Run {
    std::string general_path;
    std::string traj_file, mol_file, path_traj, path_in, path_out, prefix;
    std::string traj_type = "lmp";
    bool box_given = false, read_sequential = false, wrap = false, unwrap = false, extra_velocity = false;
    bool stream = false;          // sequential_read T and nothing in general.inp needs the whole trajectory in RAM
    bool verbose = true, custom_charges = false;
    double box_lo[3] = {0, 0, 0}, box_hi[3] = {0, 0, 0}, max_dist_sq = 0.0;
    int nsteps = 0;
    std::vector<HostType> types;
    std::map<std::string, float> mass_override, charge_override;
    int errors = 0;
    bool gpu_failed = false;      // a GPU entry point failed: pa_run_general_input then returns PA_ERR_CUDA instead of the warning count
    pa_exec exec{};
    // test seam of the speciation bookkeeping (pa_debug_run_general_input_with_connections): rows of 8 int32
    const int32_t *dbg_conns = nullptr; int64_t dbg_nconns = 0; int spec_calls = 0;
    // Distances:288-296: set_defaults() does NOT reset these two, they persist between `distance` lines
    bool dist_calc_exp = false, dist_calc_ffc = false;

    void error(int code, const std::string &msg)
    {
        ++errors;
        printf("  #  ERROR %d: %s\n", code, msg.c_str());
    }
    // the reference's NOTICEs (40, 114, 133, 140, 145, 153, 161, 163, 165, 168) take their increment of the global counter
    // back (`error_count=error_count-1`, SETTINGS:805,886,958 ...): reported, not counted
    void notice(int code, const std::string &msg) { printf("  #  NOTICE %d: %s\n", code, msg.c_str()); }
};

void set_cubic_box(Run &r, float lower, float upper)         // Molecular:931-941
{
    for (int k = 0; k < 3; ++k) { r.box_lo[k] = (double)lower; r.box_hi[k] = (double)upper; }
    const double L = r.box_hi[0] - r.box_lo[0];
    r.max_dist_sq = L * L + std::sqrt(L * L + L * L);
    r.box_given = true;
}

// first occurrence of a switch anywhere before 'quit' (Main:131-357)
const std::vector<std::string> *find_switch(const std::vector<std::vector<std::string>> &toks,
                                            std::initializer_list<const char *> names)
{
    for (const auto &t : toks) {
        if (t.empty()) continue;
        for (const char *n : names) if (t[0] == n) return &t;
        if (t[0] == "quit") return nullptr;
    }
    return nullptr;
}

bool read_molecular_input(Run &r)                             // Molecular:3444-3700
{
    std::vector<std::string> lines;
    if (!read_lines(r.mol_file, lines)) { r.error(17, "molecular input file '" + r.mol_file + "' not found"); return false; }
    std::vector<std::vector<std::string>> toks;
    for (auto &l : lines) { auto t = tokens_of(l); if (!t.empty()) toks.push_back(t); }
    int ntypes = 0;
    if (toks.size() < 2 || !parse_int(toks[0][0], r.nsteps) || !parse_int(toks[1][0], ntypes) || ntypes < 1 ||
        (int)toks.size() < 2 + ntypes) { r.error(7, "molecular input file has the wrong format"); return false; }
    r.types.assign(ntypes, HostType());
    int totalcharge = 0;
    for (int n = 0; n < ntypes; ++n) {
        HostType &t = r.types[n];
        const auto &tk = toks[2 + n];
        if (tk.size() < 3 || !parse_int(tk[0], t.charge) || !parse_int(tk[1], t.natoms) || !parse_int(tk[2], t.nmol)) {
            r.error(7, "molecular input file has the wrong format"); return false;
        }
        if (t.natoms < 1) { r.error(80, "number of atoms per molecule < 1"); return false; }
        if (t.nmol < 1) { r.error(119, "number of molecules < 1"); return false; }
        t.realcharge = (float)t.charge;
        t.elements.assign(t.natoms, "");
        t.atom_mass.assign(t.natoms, 0.0); t.atom_charge.assign(t.natoms, 0.0);
        t.mass_manual.assign(t.natoms, 0); t.charge_manual.assign(t.natoms, 0);
        totalcharge += t.charge * t.nmol;
    }
    if (totalcharge != 0) r.error(52, "total charge of the system is not zero (warning)");
    // optional sections, each searched from the line after the header up to 'quit'
    auto section = [&](std::initializer_list<const char *> names, size_t &start, int &count) -> bool {
        for (size_t i = 2 + ntypes; i < toks.size(); ++i) {
            if (toks[i][0] == "quit") return false;
            for (const char *nm : names)
                if (toks[i][0] == nm) {
                    if (toks[i].size() < 2 || !parse_int(toks[i][1], count)) return false;
                    start = i + 1;
                    return count > 0;
                }
        }
        return false;
    };
    size_t st; int cnt;
    if (section({"masses", "default_masses"}, st, cnt))
        for (int m = 0; m < cnt && st + m < toks.size(); ++m) {
            float v; const auto &tk = toks[st + m];
            if (tk.size() < 2 || !parse_real4(tk[1], v)) break;
            r.mass_override[tk[0].substr(0, 2)] = v;
        }
    if (section({"atomic_masses", "atom_masses"}, st, cnt))
        for (int m = 0; m < cnt && st + m < toks.size(); ++m) {
            int ti, ai; float v; const auto &tk = toks[st + m];
            if (tk.size() < 3 || !parse_int(tk[0], ti) || !parse_int(tk[1], ai) || !parse_real4(tk[2], v)) break;
            if (ti < 1 || ti > ntypes) { r.error(110, "invalid molecule type in atomic_masses"); continue; }
            if (ai < 1 || ai > r.types[ti - 1].natoms) { r.error(111, "invalid atom index in atomic_masses"); continue; }
            r.types[ti - 1].atom_mass[ai - 1] = (double)v; r.types[ti - 1].mass_manual[ai - 1] = 1;
        }
    if (section({"charges", "default_charges"}, st, cnt)) {
        r.custom_charges = true;
        for (int m = 0; m < cnt && st + m < toks.size(); ++m) {
            float v; const auto &tk = toks[st + m];
            if (tk.size() < 2 || !parse_real4(tk[1], v)) break;
            r.charge_override[tk[0].substr(0, 2)] = v;
        }
    }
    if (section({"atomic_charges", "atom_charges"}, st, cnt)) {
        r.custom_charges = true;
        for (int m = 0; m < cnt && st + m < toks.size(); ++m) {
            int ti, ai; float v; const auto &tk = toks[st + m];
            if (tk.size() < 3 || !parse_int(tk[0], ti) || !parse_int(tk[1], ai) || !parse_real4(tk[2], v)) break;
            if (ti < 1 || ti > ntypes) { r.error(110, "invalid molecule type in atomic_charges"); continue; }
            if (ai < 1 || ai > r.types[ti - 1].natoms) { r.error(111, "invalid atom index in atomic_charges"); continue; }
            r.types[ti - 1].atom_charge[ai - 1] = (double)v; r.types[ti - 1].charge_manual[ai - 1] = 1;
        }
    }
    return true;
}

float element_mass(Run &r, const std::string &el)
{
    auto o = r.mass_override.find(el);
    if (o != r.mass_override.end()) return o->second;
    auto d = default_masses().find(el);
    if (d != default_masses().end()) return d->second;
    r.error(4, "unknown element '" + el + "'");
    return 0.0f;
}

// Trajectory: header pass over frame 1 (elements -> masses / charges, Molecular:4827-4949) and
// body (Molecular:4951-5063): `READ(3,*) element_name, coordinates` per atom, REAL(4) items.
//
// Multi-threaded ingestion: the file is read in large blocks; each block is cut into
// whole frames by counting newlines, split among the host threads at line boundaries, and every
// thread parses its lines with parse_coordinate (decimal -> REAL(4) directly, bit-identical to strtof,
// like the list-directed READ at Molecular:5042) straight into [frame][mol][atom][3] arrays.  The
// reference's reader is serial.  TrajStream keeps the position between calls, so the same code loads
// the whole trajectory (sequential_read F) or streams it batch by batch (sequential_read T).
struct TrajStream {
    FILE *f = nullptr;
    std::vector<char> buf;
    size_t have = 0;                    // valid bytes in buf
    bool eof = false;
    size_t file_off = 0;                // bytes of the text file consumed into buf so far
    int frames_done = 0;                // frames parsed so far = index of the next frame
    double t_read = 0, t_count = 0, t_parse = 0;
    // binary frame cache (see below): read side / write side
    int cache_fd = -1;                  // >= 0: frames come from the cache, the text file is not opened
    size_t cache_data_off = 0;
    int cache_frames = 0;
    int wcache_fd = -1;                 // >= 0: parsed frames are appended to a cache under construction
    std::string wcache_tmp, wcache_final;
    ~TrajStream()
    {
        if (f) fclose(f);
        if (cache_fd >= 0) close(cache_fd);
        if (wcache_fd >= 0) { close(wcache_fd); unlink(wcache_tmp.c_str()); }      // incomplete: discard
    }
};

// Binary frame cache (optional): the first read of a text trajectory can leave the parsed REAL(4)
// coordinates next to it -- [frame][atom in file order][3] float32 after a header that pins the source file (size,
// modification time), the molecule table, the box of an lmp header and the element names -- and every later run
// (another analysis, the kernel ladder re-reading the file, several GPUs consuming frames faster than the text parser
// delivers them) reads frames with pread at memory-copy speed instead of parsing 35 bytes of text per atom.
// Enabled by the environment variable PREALPHA_B200_CACHE: "1" = next to the trajectory, otherwise a directory.
struct CacheHeader {
    char magic[8];                      // "PAB200C1"
    uint32_t version, ntypes, lmp, elem_bytes;
    uint64_t natoms_total, nsteps, src_size;
    int64_t src_mtime_ns;
    double box_lo[3], box_hi[3];
};

std::string cache_path_for(const Run &r)
{
    const char *e = getenv("PREALPHA_B200_CACHE");
    if (!e || !*e || !strcmp(e, "0")) return "";
    std::string dir = !strcmp(e, "1") ? r.path_traj : std::string(e);
    if (!dir.empty() && dir.back() != '/') dir += '/';
    std::string name = r.traj_file;
    for (char &c : name) if (c == '/') c = '_';
    return dir + name + ".pab200cache";
}

bool source_stat(const std::string &path, uint64_t &size, int64_t &mtime_ns)
{
    struct stat st;
    if (stat(path.c_str(), &st) != 0) return false;
    size = (uint64_t)st.st_size;
    mtime_ns = (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec;
    return true;
}

// opens a valid cache for this run (same source file, same molecule table, enough frames); fills elements and box
bool open_cache(Run &r, TrajStream &ts, const std::string &src)
{
    const std::string cp = cache_path_for(r);
    if (cp.empty()) return false;
    const int fd = open(cp.c_str(), O_RDONLY);
    if (fd < 0) return false;
    CacheHeader h;
    uint64_t size = 0; int64_t mt = 0;
    uint64_t natoms_total = 0;
    for (auto &t : r.types) natoms_total += (uint64_t)t.natoms * t.nmol;
    bool ok = pread(fd, &h, sizeof h, 0) == (ssize_t)sizeof h && !memcmp(h.magic, "PAB200C1", 8) && h.version == 1 &&
              source_stat(src, size, mt) && h.src_size == size && h.src_mtime_ns == mt && h.ntypes == r.types.size() &&
              h.natoms_total == natoms_total && h.lmp == (r.traj_type == "lmp" ? 1u : 0u) && h.nsteps >= (uint64_t)r.nsteps;
    std::vector<uint32_t> table(2 * r.types.size());
    std::vector<char> elems;
    if (ok) ok = pread(fd, table.data(), table.size() * 4, sizeof h) == (ssize_t)(table.size() * 4);
    if (ok) for (size_t t = 0; t < r.types.size(); ++t) if (table[2 * t] != (uint32_t)r.types[t].natoms || table[2 * t + 1] != (uint32_t)r.types[t].nmol) ok = false;
    if (ok) { elems.resize(h.elem_bytes); ok = pread(fd, elems.data(), elems.size(), sizeof h + table.size() * 4) == (ssize_t)elems.size(); }
    size_t need = 0;
    for (auto &t : r.types) need += (size_t)t.natoms * 2;
    if (ok && elems.size() != need) ok = false;
    if (!ok) { close(fd); return false; }
    size_t o = 0;
    for (auto &t : r.types) for (int a = 0; a < t.natoms; ++a, o += 2) t.elements[a] = std::string(elems.data() + o, elems[o + 1] ? 2 : (elems[o] ? 1 : 0));
    if (h.lmp) for (int k = 0; k < 3; ++k) { r.box_lo[k] = h.box_lo[k]; r.box_hi[k] = h.box_hi[k]; }
    ts.cache_fd = fd;
    ts.cache_data_off = (sizeof h + table.size() * 4 + elems.size() + 63) & ~(size_t)63;
    ts.cache_frames = (int)h.nsteps;
    return true;
}

void begin_cache_write(Run &r, TrajStream &ts)
{
    const std::string cp = cache_path_for(r);
    if (cp.empty()) return;
    ts.wcache_final = cp;
    ts.wcache_tmp = cp + ".tmp." + std::to_string((long long)getpid());
    ts.wcache_fd = open(ts.wcache_tmp.c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0644);
    if (ts.wcache_fd < 0) return;
    size_t elem_bytes = 0;
    for (auto &t : r.types) elem_bytes += (size_t)t.natoms * 2;
    ts.cache_data_off = (sizeof(CacheHeader) + r.types.size() * 8 + elem_bytes + 63) & ~(size_t)63;
}

// the header is written last (elements and the lmp box are known after frame 1, the frame count at the end); the file
// becomes visible under its final name only when it is complete
void finish_cache_write(Run &r, TrajStream &ts, const std::string &src)
{
    if (ts.wcache_fd < 0) return;
    CacheHeader h{};
    memcpy(h.magic, "PAB200C1", 8);
    h.version = 1; h.ntypes = (uint32_t)r.types.size(); h.lmp = r.traj_type == "lmp" ? 1u : 0u;
    h.natoms_total = 0;
    for (auto &t : r.types) h.natoms_total += (uint64_t)t.natoms * t.nmol;
    h.nsteps = (uint64_t)ts.frames_done;
    bool ok = source_stat(src, h.src_size, h.src_mtime_ns);
    for (int k = 0; k < 3; ++k) { h.box_lo[k] = r.box_lo[k]; h.box_hi[k] = r.box_hi[k]; }
    std::vector<uint32_t> table;
    std::vector<char> elems;
    for (auto &t : r.types) {
        table.push_back((uint32_t)t.natoms); table.push_back((uint32_t)t.nmol);
        for (int a = 0; a < t.natoms; ++a) { elems.push_back(t.elements[a].size() > 0 ? t.elements[a][0] : 0); elems.push_back(t.elements[a].size() > 1 ? t.elements[a][1] : 0); }
    }
    h.elem_bytes = (uint32_t)elems.size();
    ok = ok && pwrite(ts.wcache_fd, &h, sizeof h, 0) == (ssize_t)sizeof h &&
         pwrite(ts.wcache_fd, table.data(), table.size() * 4, sizeof h) == (ssize_t)(table.size() * 4) &&
         pwrite(ts.wcache_fd, elems.data(), elems.size(), sizeof h + table.size() * 4) == (ssize_t)elems.size();
    close(ts.wcache_fd);
    ts.wcache_fd = -1;
    if (ok && rename(ts.wcache_tmp.c_str(), ts.wcache_final.c_str()) == 0) {
        if (r.verbose) printf(" wrote the binary frame cache '%s' (%d steps).\n", ts.wcache_final.c_str(), ts.frames_done);
    } else unlink(ts.wcache_tmp.c_str());
}

bool open_stream(Run &r, TrajStream &ts, bool may_write_cache = true)
{
    const std::string path = r.path_traj + r.traj_file;
    if (open_cache(r, ts, path)) return true;
    ts.f = fopen(path.c_str(), "r");
    if (!ts.f) { r.error(9, "trajectory file '" + path + "' not found"); return false; }
    if (may_write_cache) begin_cache_write(r, ts);
    return true;
}

// Parses up to `want` further frames into dst[type] ([local frame][mol][atom][3], local frame 0 = first frame of this
// call); returns the number of frames parsed (< want at the end of the file).
int read_frames(Run &r, TrajStream &ts, int want_frames, const std::vector<float *> &dst)
{
    auto now_s = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const bool lmp = (r.traj_type == "lmp");
    const int header = lmp ? 9 : 2;
    long long natoms_total = 0;
    for (auto &t : r.types) natoms_total += (long long)t.natoms * t.nmol;
    const long long lines_per_frame = header + natoms_total;
    std::vector<long long> type_first(r.types.size() + 1, 0);           // first atom line of each type inside a frame
    for (size_t t = 0; t < r.types.size(); ++t) type_first[t + 1] = type_first[t] + (long long)r.types[t].natoms * r.types[t].nmol;
    if (ts.cache_fd >= 0) {
        // binary cache: a frame is natoms_total * 12 bytes in file order, i.e. one contiguous block per molecule type
        const double t0 = now_s();
        const int n = std::max(0, std::min(want_frames, ts.cache_frames - ts.frames_done));
        const unsigned nth = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        std::vector<std::thread> th;
        bool bad = false;
        for (unsigned w = 0; w < nth; ++w)
            th.emplace_back([&, w] {
                for (int s = (int)w; s < n; s += (int)nth)
                    for (size_t t = 0; t < r.types.size(); ++t) {
                        const size_t bytes = (size_t)(type_first[t + 1] - type_first[t]) * 12;
                        const off_t off = (off_t)(ts.cache_data_off + ((size_t)(ts.frames_done + s) * (size_t)natoms_total + (size_t)type_first[t]) * 12);
                        char *out = reinterpret_cast<char *>(dst[t] + (size_t)s * (size_t)(type_first[t + 1] - type_first[t]) * 3);
                        size_t done = 0;
                        while (done < bytes) {
                            const ssize_t got = pread(ts.cache_fd, out + done, bytes - done, off + (off_t)done);
                            if (got <= 0) { bad = true; break; }
                            done += (size_t)got;
                        }
                    }
            });
        for (auto &x : th) x.join();
        ts.t_read += now_s() - t0;
        if (bad) return 0;
        ts.frames_done += n;
        return n;
    }
    auto parse_atom_line = [&](const char *p, int step, long long a_total, bool first_frame) {
        size_t t = 0;
        while (t + 1 < r.types.size() && a_total >= type_first[t + 1]) ++t;
        HostType &ty = r.types[t];
        const long long local = a_total - type_first[t];
        const int m = (int)(local / ty.natoms), a = (int)(local % ty.natoms);
        while (*p == ' ' || *p == '\t') ++p;
        const char *e = p;
        while (*e && *e != ' ' && *e != '\t' && *e != ',' && *e != '\n' && *e != '\r') ++e;
        if (first_frame && m == 0) ty.elements[a] = std::string(p, (size_t)std::min<ptrdiff_t>(e - p, 2));
        float *out = dst[t] + (((size_t)step * ty.nmol + m) * ty.natoms + a) * 3;
        p = e;
        for (int k = 0; k < 3; ++k) {
            while (*p == ' ' || *p == '\t' || *p == ',') ++p;
            const char *q; out[k] = parse_coordinate(p, &q); p = q;
        }
    };
    auto parse_header_line = [&](const std::string &ln, int h) {        // frame 1 only, serial
        if (lmp && h >= 5 && h <= 7) {                                    // box bounds, Molecular:4851-4857
            auto tk = tokens_of(ln);
            if (tk.size() < 2 || !parse_real8(tk[0], r.box_lo[h - 5]) || !parse_real8(tk[1], r.box_hi[h - 5])) r.error(71, "cannot read box from lammps header");
        }
        if (lmp && h == 8) {
            auto tk = tokens_of(ln);
            if (tk.size() < 6 || tk[2] != "element") r.error(53, "lammps trajectory must have 'element' as first column");
            else if (!(tk[3] == "xu" && tk[4] == "yu" && tk[5] == "zu")) r.error(54, "expected 'xu yu zu' columns");
        }
        if (h == (lmp ? 3 : 0)) {
            auto tk = tokens_of(ln);
            int n = 0;
            if (tk.empty() || !parse_int(tk[0], n)) r.error(71, "cannot read number of atoms");
            else if (n != natoms_total) r.error(10, "number of atoms in trajectory and molecular input differ");
        }
    };
    const unsigned nthreads = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
    size_t block_bytes = (size_t)64 << 20;
    if (const char *e = getenv("PA_READ_BLOCK")) block_bytes = (size_t)std::max(64, atoi(e));      // test hook: bytes per read
    std::vector<char> &buf = ts.buf;
    int steps_read = 0;
    // runs fn(k, begin, end) for nthreads equal byte ranges of [0, n) on the host threads (ranges >= `grain` bytes)
    auto parallel_ranges = [&](size_t n, size_t grain, const std::function<void(unsigned, size_t, size_t)> &fn) {
        const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(nthreads, n / std::max<size_t>(grain, 1)));
        if (nt <= 1) { fn(0, 0, n); return 1u; }
        std::vector<std::thread> th;
        for (unsigned k = 0; k < nt; ++k) th.emplace_back([&, k] { fn(k, n * k / nt, n * (k + 1) / nt); });
        for (auto &x : th) x.join();
        return nt;
    };
    while (steps_read < want_frames) {
        // 1. top up the buffer (only when it does not already hold a whole frame)
        double t0 = now_s();
        const int want = want_frames - steps_read;
        long long nl = 0; size_t end = 0; int nframes = 0;
        // how many whole frames does the buffer hold, and where does the last of them end?  Newlines are counted on all
        // host threads (byte ranges), the frame boundary is then located inside the one range that contains it.
        auto count_frames = [&] {
            nl = 0; end = 0; nframes = 0;
            std::vector<long long> cnt(nthreads, 0);
            std::vector<size_t> rb(nthreads + 1, ts.have);
            const unsigned nt = parallel_ranges(ts.have, (size_t)1 << 20, [&](unsigned k, size_t b, size_t e) {
                rb[k] = b;
                long long c = 0;
                for (const char *p = buf.data() + b, *lim = buf.data() + e; p < lim;) {
                    const char *q = (const char *)memchr(p, '\n', (size_t)(lim - p));
                    if (!q) break;
                    ++c; p = q + 1;
                }
                cnt[k] = c;
            });
            rb[nt] = ts.have;
            for (unsigned k = 0; k < nt; ++k) nl += cnt[k];
            nframes = (int)std::min<long long>(want, nl / lines_per_frame);
            if (nframes == 0) return;
            long long need = (long long)nframes * lines_per_frame;         // the need-th newline ends the last whole frame
            unsigned k = 0;
            while (k + 1 < nt && need > cnt[k]) { need -= cnt[k]; ++k; }
            const char *p = buf.data() + rb[k], *lim = buf.data() + rb[k + 1];
            while (need > 0 && p < lim) {
                const char *q = (const char *)memchr(p, '\n', (size_t)(lim - p));
                if (!q) break;
                --need; p = q + 1;
            }
            end = (size_t)(p - buf.data());
            nl = (long long)nframes * lines_per_frame;
        };
        if (ts.have > 0) count_frames();
        ts.t_count += now_s() - t0; t0 = now_s();
        if (nframes < want && !ts.eof) {
            if (buf.size() < ts.have + block_bytes + 1) buf.resize(ts.have + block_bytes + 1);
            // the next block of the file, read by several threads at once (page-cache copies scale with the threads)
            const int fd = fileno(ts.f);
            std::vector<size_t> part(nthreads, 0);
            std::vector<size_t> asked(nthreads, 0);
            const off_t base = (off_t)ts.file_off;
            char *into = buf.data() + ts.have;
            const unsigned nt = parallel_ranges(block_bytes, (size_t)4 << 20, [&](unsigned k, size_t b, size_t e) {
                asked[k] = e - b;
                size_t done = 0;
                while (done < e - b) {
                    const ssize_t got = pread(fd, into + b + done, e - b - done, base + (off_t)(b + done));
                    if (got <= 0) break;
                    done += (size_t)got;
                }
                part[k] = done;
            });
            size_t got = 0;                                               // contiguous bytes from the start of the block
            for (unsigned k = 0; k < nt; ++k) { got += part[k]; if (part[k] < asked[k]) break; }
            ts.file_off += got;
            ts.have += got;
            if (got < block_bytes) { ts.eof = true; if (ts.have && buf[ts.have - 1] != '\n') buf[ts.have++] = '\n'; }
            ts.t_read += now_s() - t0; t0 = now_s();
            // 2. how many whole frames does it hold now?
            count_frames();
            ts.t_count += now_s() - t0; t0 = now_s();
        }
        if (nframes == 0) {
            if (ts.eof) break;                                            // truncated file: the caller reports error 84
            continue;                                                     // frame larger than the buffer so far: read more
        }
        // 3. very first frame of the file: header lines serially
        if (ts.frames_done == 0) {
            const char *p = buf.data();
            for (int h = 0; h < header; ++h) {
                const char *q = (const char *)memchr(p, '\n', end - (size_t)(p - buf.data()));
                parse_header_line(std::string(p, (size_t)(q - p)), h);
                p = q + 1;
            }
        }
        // 4. split [0,end) among the threads at line boundaries, count lines, then parse
        std::vector<size_t> cut(nthreads + 1, end);
        cut[0] = 0;
        for (unsigned t = 1; t < nthreads; ++t) {
            size_t pos = end / nthreads * t;
            const char *q = (const char *)memchr(buf.data() + pos, '\n', end - pos);
            cut[t] = q ? (size_t)(q - buf.data()) + 1 : end;
        }
        for (unsigned t = 1; t <= nthreads; ++t) if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
        std::vector<long long> nlines(nthreads, 0);
        {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nthreads; ++t)
                th.emplace_back([&, t] {
                    long long c = 0;
                    for (const char *p = buf.data() + cut[t], *lim = buf.data() + cut[t + 1]; p < lim;) {
                        const char *q = (const char *)memchr(p, '\n', (size_t)(lim - p));
                        if (!q) break;
                        ++c; p = q + 1;
                    }
                    nlines[t] = c;
                });
            for (auto &x : th) x.join();
        }
        std::vector<long long> first_line(nthreads + 1, 0);
        for (unsigned t = 0; t < nthreads; ++t) first_line[t + 1] = first_line[t] + nlines[t];
        {
            std::vector<std::thread> th;
            const int step0 = steps_read;
            const bool file_start = ts.frames_done == 0;
            for (unsigned t = 0; t < nthreads; ++t)
                th.emplace_back([&, t] {
                    long long ln = first_line[t];
                    for (const char *p = buf.data() + cut[t], *lim = buf.data() + cut[t + 1]; p < lim; ++ln) {
                        const char *q = (const char *)memchr(p, '\n', (size_t)(lim - p));
                        if (!q) break;
                        const long long in_frame = ln % lines_per_frame;
                        const int step = step0 + (int)(ln / lines_per_frame);
                        if (in_frame >= header) parse_atom_line(p, step, in_frame - header, file_start && step == 0);
                        p = q + 1;
                    }
                });
            for (auto &x : th) x.join();
        }
        ts.t_parse += now_s() - t0;
        if (ts.wcache_fd >= 0) {                                            // append the parsed frames to the cache under construction
            for (int s2 = 0; s2 < nframes && ts.wcache_fd >= 0; ++s2)
                for (size_t t = 0; t < r.types.size(); ++t) {
                    const size_t bytes = (size_t)(type_first[t + 1] - type_first[t]) * 12;
                    const off_t off = (off_t)(ts.cache_data_off + ((size_t)(ts.frames_done + s2) * (size_t)natoms_total + (size_t)type_first[t]) * 12);
                    const char *src = reinterpret_cast<const char *>(dst[t] + (size_t)(steps_read + s2) * (size_t)(type_first[t + 1] - type_first[t]) * 3);
                    if (pwrite(ts.wcache_fd, src, bytes, off) != (ssize_t)bytes) { close(ts.wcache_fd); ts.wcache_fd = -1; unlink(ts.wcache_tmp.c_str()); break; }
                }
        }
        steps_read += nframes;
        ts.frames_done += nframes;
        // 5. keep the unparsed tail
        memmove(buf.data(), buf.data() + end, ts.have - end);
        ts.have -= end;
    }
    return steps_read;
}

// sequential_read F: the whole trajectory; sequential_read T: frame 1 only (elements, masses, box) -- the analyses
// then stream the file themselves (perform_distribution_streaming).
bool load_trajectory(Run &r)
{
    TrajStream ts;
    if (!open_stream(r, ts, /*may_write_cache=*/!r.stream)) return false;
    if (ts.cache_fd >= 0 && r.verbose) printf(" reading the binary frame cache '%s' instead of the text file.\n", cache_path_for(r).c_str());
    const auto t_load0 = std::chrono::steady_clock::now();
    const bool lmp = (r.traj_type == "lmp");
    long long natoms_total = 0;
    const int want = r.stream ? std::min(r.nsteps, 1) : r.nsteps;
    std::vector<float *> dst;
    for (auto &t : r.types) { natoms_total += (long long)t.natoms * t.nmol; t.xyz.assign((size_t)want * t.natoms * t.nmol * 3, 0.f); dst.push_back(t.xyz.data()); }
    const int steps_read = read_frames(r, ts, want, dst);
    if (!r.stream && steps_read > 0) finish_cache_write(r, ts, r.path_traj + r.traj_file);
    if (getenv("PA_DEBUG_TIMING")) fprintf(stderr, "[prealpha_b200] ingestion: read %.3f s, frame cut %.3f s, parse %.3f s\n", ts.t_read, ts.t_count, ts.t_parse);
    if (steps_read < want) {                                      // Molecular:5003-5009, error 84
        r.error(84, "trajectory is shorter than specified; using " + std::to_string(steps_read) + " steps");
        // re-pack is not needed: arrays are [step][mol][atom][3], truncation just drops the tail
        for (auto &t : r.types) t.xyz.resize((size_t)steps_read * t.natoms * t.nmol * 3);
        r.nsteps = steps_read;
    }
    if (lmp) {
        r.box_given = true;
        const double Lx = r.box_hi[0] - r.box_lo[0], Ly = r.box_hi[1] - r.box_lo[1], Lz = r.box_hi[2] - r.box_lo[2];
        r.max_dist_sq = Ly * Ly + std::sqrt(Lx * Lx + Lz * Lz);   // Molecular:4857 [sic]
    } else {
        r.max_dist_sq = 22500.0;                                   // Molecular:4883
    }
    // masses / charges from the elements of frame 1 (Molecular:4896-4924)
    for (auto &t : r.types) {
        t.mass = 0.0;
        for (int a = 0; a < t.natoms; ++a) {
            float em = element_mass(r, t.elements[a]);
            if (t.mass_manual[a]) em = (float)t.atom_mass[a]; else t.atom_mass[a] = (double)em;
            float ec;
            if (t.natoms == 1) ec = (float)t.charge;
            else { auto o = r.charge_override.find(t.elements[a]); ec = o != r.charge_override.end() ? o->second : 0.0f; }
            if (!t.charge_manual[a]) t.atom_charge[a] = (double)ec;
            t.mass = t.mass + (double)em;
        }
        if (r.custom_charges) {                                    // Molecular:4712-4729
            double sum = 0.0;
            for (int a = 0; a < t.natoms; ++a) sum = sum + t.atom_charge[a];
            if (std::fabs(sum - (double)(float)t.charge) > (double)0.001f) { r.error(126, "sum of atomic charges differs from formal charge"); t.realcharge = (float)sum; }
        }
    }
    if (r.verbose) {
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_load0).count();
        printf(" reading the trajectory...done (%d steps of %lld atoms in %.3f s: %.1f million atoms/s on the host threads).\n",
               steps_read, natoms_total, sec, sec > 0 ? (double)steps_read * (double)natoms_total / sec * 1e-6 : 0.0);
    }
    return true;
}


// ---- wrap_trajectory / unwrap_trajectory pre-passes ---------------------------------------
// O(N) per frame scalar work that rewrites the stored REAL(4) coordinates before any analysis
// (Molecular:4801-4816).  Host threads; molecules are independent.

void host_center_of_mass(const Run &r, const HostType &ty, int step, int mol, bool com_boost, double c[3])   // Molecular:2804-2834
{
    const float *p = &ty.xyz[((size_t)step * ty.nmol + mol) * (size_t)ty.natoms * 3];
    if (com_boost) { c[0] = (double)p[0]; c[1] = (double)p[1]; c[2] = (double)p[2]; return; }
    (void)r;
    c[0] = c[1] = c[2] = 0.0;
    for (int a = 0; a < ty.natoms; ++a)
        for (int k = 0; k < 3; ++k) { double w = (double)p[a * 3 + k]; w = w * ty.atom_mass[a]; c[k] = c[k] + w; }
    for (int k = 0; k < 3; ++k) c[k] = c[k] / ty.mass;
}

template <typename F>
void parallel_for(int n, F fn)
{
    const unsigned nt = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
    if (n < 2 || nt < 2) { for (int i = 0; i < n; ++i) fn(i); return; }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back([&, t] { for (int i = (int)t; i < n; i += (int)nt) fn(i); });
    for (auto &x : th) x.join();
}

bool all_one_atom(const Run &r) { for (auto &t : r.types) if (t.natoms > 1) return false; return true; }

void wrap_full(Run &r)                                          // Molecular:1703-1765
{
    const bool boost = all_one_atom(r);
    parallel_for(r.nsteps, [&](int step) {
        for (auto &ty : r.types)
            for (int m = 0; m < ty.nmol; ++m) {
                double c[3];
                host_center_of_mass(r, ty, step, m, boost, c);
                float *p = &ty.xyz[((size_t)step * ty.nmol + m) * (size_t)ty.natoms * 3];
                for (int k = 0; k < 3; ++k) {
                    float com = (float)c[k];                                 // REAL :: centre_of_mass(3)
                    const double L = r.box_hi[k] - r.box_lo[k];
                    for (int it = 0; it < 1000; ++it) {
                        if ((double)com < r.box_lo[k]) {
                            com = (float)((double)com + L);
                            for (int a = 0; a < ty.natoms; ++a) p[a * 3 + k] = (float)((double)p[a * 3 + k] + L);
                        } else if ((double)com > r.box_hi[k]) {
                            com = (float)((double)com - L);
                            for (int a = 0; a < ty.natoms; ++a) p[a * 3 + k] = (float)((double)p[a * 3 + k] - L);
                        } else break;
                    }
                }
            }
    });
}

// give_smallest_distance_squared with translation (Molecular:1430-1469), WRAP_TRAJECTORY = F
void host_link_vector(const Run &r, const double c1[3], const double c2[3], double tr[3])
{
    double p1[3], p2[3], sh1[3], sh2[3], L[3];
    for (int k = 0; k < 3; ++k) {
        L[k] = r.box_hi[k] - r.box_lo[k];
        p1[k] = c1[k]; p2[k] = c2[k]; sh1[k] = sh2[k] = 0.0;
        for (int it = 0; it < 1000; ++it) { if (p1[k] < r.box_lo[k]) { p1[k] += L[k]; sh1[k] += L[k]; } else if (p1[k] > r.box_hi[k]) { p1[k] -= L[k]; sh1[k] -= L[k]; } else break; }
        for (int it = 0; it < 1000; ++it) { if (p2[k] < r.box_lo[k]) { p2[k] += L[k]; sh2[k] += L[k]; } else if (p2[k] > r.box_hi[k]) { p2[k] -= L[k]; sh2[k] -= L[k]; } else break; }
    }
    double best = r.max_dist_sq;
    tr[0] = tr[1] = tr[2] = 0.0;
    for (int a = -1; a <= 1; ++a) for (int b = -1; b <= 1; ++b) for (int c = -1; c <= 1; ++c) {
        const double s[3] = { (double)a * L[0], (double)b * L[1], (double)c * L[2] };
        const double d0 = (p2[0] + s[0]) - p1[0], d1 = (p2[1] + s[1]) - p1[1], d2 = (p2[2] + s[2]) - p1[2];
        const double clip = (d0 * d0 + d1 * d1) + d2 * d2;
        if (clip < best) { best = clip; tr[0] = s[0]; tr[1] = s[1]; tr[2] = s[2]; }
    }
    for (int k = 0; k < 3; ++k) tr[k] = tr[k] + (sh2[k] - sh1[k]);
}

void unwrap_full(Run &r)                                        // Molecular:1590-1699
{
    const bool boost = all_one_atom(r);
    double thr[3];
    for (int k = 0; k < 3; ++k) thr[k] = (r.box_hi[k] - r.box_lo[k]) / 2.0;
    for (auto &ty : r.types)
        parallel_for(ty.nmol, [&](int m) {
            double acc[3] = { 0, 0, 0 };
            bool jumped = false;
            for (int step = 0; step + 1 < r.nsteps; ++step) {
                float *pn = &ty.xyz[((size_t)(step + 1) * ty.nmol + m) * (size_t)ty.natoms * 3];
                if (jumped) {
                    if ((acc[0] * acc[0] + acc[1] * acc[1]) + acc[2] * acc[2] < 0.001) jumped = false;
                    else for (int a = 0; a < ty.natoms; ++a) for (int k = 0; k < 3; ++k) pn[a * 3 + k] = (float)((double)pn[a * 3 + k] + acc[k]);
                }
                double c1[3], c2[3], link[3];
                host_center_of_mass(r, ty, step, m, boost, c1);
                host_center_of_mass(r, ty, step + 1, m, boost, c2);
                host_link_vector(r, c1, c2, link);
                if (std::fabs(link[0]) > thr[0] || std::fabs(link[1]) > thr[1] || std::fabs(link[2]) > thr[2]) {
                    for (int a = 0; a < ty.natoms; ++a) for (int k = 0; k < 3; ++k) pn[a * 3 + k] = (float)((double)pn[a * 3 + k] + link[k]);
                    jumped = true;
                    for (int k = 0; k < 3; ++k) acc[k] = acc[k] + link[k];
                }
            }
        });
}

struct DistInput {
    int mode = PA_MODE_PDF;
    std::vector<std::array<int, 5>> refs;
    int sampling_interval = 10, bin_a = 100, bin_b = 100;        // Distribution:9-16
    bool subtract_uniform = false, weigh_charge = false, use_coc = false, maxdist_optimize = false, normalise_clm = false;
    float maxdist = 10.0f;
};

int read_distribution_input(Run &r, const std::string &name, DistInput &d)
{
    std::vector<std::string> lines;
    if (!read_lines(r.path_in + name, lines)) { r.error(132, "distribution input file '" + r.path_in + name + "' not found"); return 132; }
    std::vector<std::vector<std::string>> toks;
    for (auto &l : lines) { auto t = tokens_of(l); if (!t.empty()) toks.push_back(t); }
    if (toks.empty()) { r.error(99, "incorrect format in distribution input"); return 99; }
    std::string mode = toks[0][0];
    if (mode == "cydf") mode = "cdf";
    if (mode == "podf") mode = "pdf";
    if (mode == "cdf") d.mode = PA_MODE_CDF;
    else if (mode == "pdf") d.mode = PA_MODE_PDF;
    else if (mode == "charge_arm") d.mode = PA_MODE_CHARGE_ARM;
    else { r.error(99, "incorrect format in distribution input"); return 99; }
    int nref = 0;
    if (toks[0].size() < 2 || !parse_int(toks[0][1], nref) || nref < 0 || (int)toks.size() < 1 + nref) { r.error(99, "incorrect format in distribution input"); return 99; }
    const int ntypes = (int)r.types.size();
    for (int n = 0; n < nref; ++n) {
        std::array<int, 5> v{};
        const auto &tk = toks[1 + n];
        const int nfields = d.mode == PA_MODE_CHARGE_ARM ? 4 : 5;            // Distribution:427-445: charge_arm references have no observed type
        bool ok = (int)tk.size() >= nfields;
        for (int k = 0; ok && k < nfields; ++k) ok = parse_int(tk[k], v[k]);
        if (!ok) { r.error(99, "incorrect format in distribution input"); return 99; }
        if (v[3] > ntypes || v[3] < 1 || v[4] > ntypes) { r.error(33, "molecule type index does not exist"); return 33; }
        d.refs.push_back(v);
    }
    for (size_t i = 1 + nref; i < toks.size() && i < (size_t)(1 + nref + 500); ++i) {
        const auto &tk = toks[i];
        const std::string &k = tk[0];
        auto want_int = [&](int &dst, int dflt) { int v; if (tk.size() > 1 && parse_int(tk[1], v)) dst = v; else { r.error(100, "cannot read '" + k + "'"); dst = dflt; } };
        auto want_log = [&](bool &dst, bool dflt) { bool v; if (tk.size() > 1 && parse_logical(tk[1], v)) dst = v; else { r.error(100, "cannot read '" + k + "'"); dst = dflt; } };
        if (k == "sampling_interval") want_int(d.sampling_interval, 10);
        else if (k == "bin_count") { int v = 100; want_int(v, 100); d.bin_a = d.bin_b = v; }
        else if (k == "bin_count_a") want_int(d.bin_a, 100);
        else if (k == "bin_count_b") want_int(d.bin_b, 100);
        else if (k == "maxdist") {
            d.maxdist_optimize = false;
            float v; if (tk.size() > 1 && parse_real4(tk[1], v)) d.maxdist = v; else { r.error(100, "cannot read 'maxdist'"); d.maxdist = 10.0f; }
        } else if (k == "maxdist_optimize" || k == "maxdist_optimise") d.maxdist_optimize = true;
        else if (k == "subtract_uniform") {
            if (d.mode != PA_MODE_CDF) want_log(d.subtract_uniform, false);
            else printf("   subtract_uniform is only available for polar distribution functions.\n");
        } else if (k == "weigh_charge") {
            if (d.mode == PA_MODE_CHARGE_ARM) printf(" charge weighing not available for charge_arm analysis.\n");
            else want_log(d.weigh_charge, false);
        } else if (k == "normalise_CLM" || k == "normalise_charge_arm" || k == "normalize_CLM" || k == "normalize_charge_arm" || k == "normalise_clm") {
            if (d.mode == PA_MODE_CHARGE_ARM) want_log(d.normalise_clm, false);
            else printf(" The charge lever moment normalisation is only available for charge_arm analysis.\n");
        } else if (k == "use_coc" || k == "use_COC" || k == "center_of_charge" || k == "centre_of_charge") {
            if (d.mode == PA_MODE_CHARGE_ARM) printf(" center_of_charge not available for charge_arm analysis.\n");
            else want_log(d.use_coc, false);
        }
        else if (k == "quit") break;
        if (d.bin_a < 1) { d.bin_a = 1; r.error(104, "bin count out of range"); } else if (d.bin_a > 1000) { d.bin_a = 1000; r.error(104, "bin count out of range"); }
        if (d.bin_b < 1) { d.bin_b = 1; r.error(104, "bin count out of range"); } else if (d.bin_b > 1000) { d.bin_b = 1000; r.error(104, "bin count out of range"); }
    }
    return 0;
}

void make_traj(Run &r, std::vector<pa_type> &types, pa_traj &t)
{
    types.resize(r.types.size());
    bool boost = true;
    for (size_t i = 0; i < r.types.size(); ++i) {
        HostType &h = r.types[i];
        pa_type &p = types[i];
        memset(&p, 0, sizeof p);
        p.charge = h.charge; p.natoms = h.natoms; p.nmol = h.nmol; p.mass = h.mass; p.realcharge = h.realcharge;
        p.atom_mass = h.atom_mass.data(); p.atom_charge = h.atom_charge.data(); p.xyz = h.xyz.data();
        if (h.natoms > 1) boost = false;
    }
    memset(&t, 0, sizeof t);
    t.ntypes = (int)types.size(); t.nsteps = r.nsteps; t.types = types.data();
    for (int k = 0; k < 3; ++k) { t.box_lo[k] = r.box_lo[k]; t.box_hi[k] = r.box_hi[k]; }
    t.max_dist_sq = r.max_dist_sq;
    t.com_boost = boost ? 1 : 0;                              // Molecular:3556-3562
    t.wrap_trajectory = r.wrap ? 1 : 0;
}

// sequential_read T: the trajectory never sits in host memory as a whole.  Batches of frames are parsed from the file
// (read_frames: text on all host threads, or pread from the binary cache) into a staging array, handed to a session
// (pinned copy + cudaMemcpyAsync on its copy stream) and evaluated while the host already reads the next batch; the
// sampled frames of a batch are picked by the stride of pa_session_upload.  With several devices (pa_exec.ndevices) one
// worker thread per GPU owns a session and takes the next batch the reading thread has queued as soon as it can, and the
// int64 accumulators are combined on the
// first device at the end (pa_sessions_reduce_to_first: peer copies or ncclReduce over NVLink, by size).  The kernel ladder of
// pa_distribution_histogram_multi (cells -> brute force -> general kernel when a sticky device flag asks for it) is
// repeated here by re-reading the file.
int histogram_streaming(Run &r, const pa_traj &meta, std::vector<pa_job> &jobs, std::vector<int64_t *> &hp,
                        std::vector<int64_t> &within, std::vector<int64_t> &oob, pa_stats *st, pa_exec ex)
{
    const int dt = jobs[0].sampling_interval;
    long long floats_per_frame = 0;
    for (auto &t : r.types) floats_per_frame += (long long)t.natoms * t.nmol * 3;
    int ndev = ex.ndevices < 0 ? pa_device_count() - ex.device : std::max(1, ex.ndevices);
    if (ndev < 1 || ex.device + ndev > pa_device_count()) { pa::set_error("pa_exec.device/ndevices: not that many usable devices"); return PA_ERR_BAD_ARG; }
    ex.ndevices = 1;
    // frames per batch: ~32 MB of raw coordinates (two page-locked staging buffers of that size per session: page-locking
    // is slow, and with several devices the allocations serialise in the driver), at least one frame
    int B = (int)std::max<long long>(1, std::min<long long>(256, ((long long)32 << 20) / std::max<long long>(1, floats_per_frame * 4)));
    B = (int)std::min<long long>(B, std::max<long long>(1, ((long long)r.nsteps + 8 * ndev - 1) / (8 * ndev)));   // at least ~8 batches per device: reading overlaps the kernels
    if (const char *e = getenv("PA_STREAM_BATCH")) B = std::max(1, std::min(256, atoi(e)));      // test hook: frames per batch
    const std::string src = r.path_traj + r.traj_file;
    for (int attempt = 0; attempt < 3; ++attempt) {
        TrajStream ts;
        if (!open_stream(r, ts)) return PA_ERR_IO;
        std::vector<pa_session *> sess(ndev, nullptr);
        int rc = 0;
        // NCCL communicators take a few hundred ms to create: do it while the file is read and the kernels run
        std::thread comm_thread([&] { if (ndev > 1) pa_multi_device_prepare(ex.device, ndev, jobs.data(), (int)jobs.size()); });
        struct CommJoin { std::thread &t; ~CommJoin() { if (t.joinable()) t.join(); } } comm_join{ comm_thread };
        // ---- producer / consumers -------------------------------------------------------------------------------
        // This thread reads batches into a ring of host buffers; one thread per device creates its session and then
        // takes whatever batch is next (pinned copy + async upload + launches), so a device that is ready early, or
        // faster, simply takes more batches.  The sums are integers: the dealing order cannot change the result.
        struct Slot { std::vector<std::vector<float>> xyz; std::vector<float *> dst; std::vector<pa_type> vt; pa_traj view; int n = 0, g0 = 0; };
        const int nslots = ndev + 2;
        std::vector<Slot> slots(nslots);
        for (Slot &sl : slots) {
            sl.xyz.resize(r.types.size());
            sl.vt.assign(meta.types, meta.types + meta.ntypes);
            for (size_t t = 0; t < r.types.size(); ++t) {
                sl.xyz[t].assign((size_t)B * r.types[t].natoms * r.types[t].nmol * 3, 0.f);
                sl.dst.push_back(sl.xyz[t].data());
                sl.vt[t].xyz = sl.xyz[t].data();
            }
            sl.view = meta;
            sl.view.types = sl.vt.data();
        }
        std::mutex mu;
        std::condition_variable cv;
        std::deque<int> free_slots, full_slots;
        for (int i = 0; i < nslots; ++i) free_slots.push_back(i);
        bool done = false;                                        // no more batches will be queued
        int first_rc = 0;                                         // first failure of any thread (stops everybody)
        std::string first_err;
        const bool dbg_t = getenv("PA_DEBUG_TIMING") != nullptr;
        auto now_s = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        std::vector<double> t_create(ndev, 0.0), t_upload(ndev, 0.0), t_run(ndev, 0.0);
        std::vector<long long> dealt(ndev, 0);
        auto fail = [&](int code, const std::string &msg) {
            std::lock_guard<std::mutex> lk(mu);
            if (!first_rc) { first_rc = code; first_err = msg; }
            cv.notify_all();
        };
        std::vector<std::thread> workers;
        for (int d = 0; d < ndev; ++d)
            workers.emplace_back([&, d] {
                pa_exec e = ex;
                e.device = ex.device + d;
                double t0 = now_s();
                int wrc = pa_session_create(&meta, jobs.data(), (int)jobs.size(), &e, (B + dt - 1) / dt, &sess[d]);
                t_create[d] = now_s() - t0;
                if (wrc) { fail(wrc, pa_last_error()); return; }
                for (;;) {
                    int i;
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [&] { return first_rc || done || !full_slots.empty(); });
                        if (first_rc || full_slots.empty()) return;
                        i = full_slots.front();
                        full_slots.pop_front();
                    }
                    Slot &sl = slots[i];
                    const int first = (dt - sl.g0 % dt) % dt;     // sampled frames are the global multiples of dt (Distribution:718)
                    sl.view.nsteps = sl.n;
                    t0 = now_s();
                    wrc = pa_session_upload(sess[d], &sl.view, first, (sl.n - first + dt - 1) / dt, dt);
                    t_upload[d] += now_s() - t0;
                    {                                             // the frames sit in the session's pinned buffer now
                        std::lock_guard<std::mutex> lk(mu);
                        free_slots.push_back(i);
                        cv.notify_all();
                    }
                    t0 = now_s();
                    if (wrc == 0) wrc = pa_session_run(sess[d]);
                    t_run[d] += now_s() - t0;
                    ++dealt[d];
                    if (wrc) { fail(wrc, pa_last_error()); return; }
                }
            });
        int g0 = 0;                                               // global index of the first frame of the batch
        long long nbatch = 0;
        bool short_file = false;
        double t_read = 0, t_wait = 0;
        const double t_loop0 = now_s();
        while (g0 < r.nsteps) {
            int i;
            {
                double t0 = now_s();
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return first_rc || !free_slots.empty(); });
                if (first_rc) break;
                i = free_slots.front();
                free_slots.pop_front();
                t_wait += now_s() - t0;
            }
            const int want = std::min(B, r.nsteps - g0);
            const double t0 = now_s();
            const int n = read_frames(r, ts, want, slots[i].dst);
            t_read += now_s() - t0;
            const bool sampled = n > 0 && (dt - g0 % dt) % dt < n;
            {
                std::lock_guard<std::mutex> lk(mu);
                if (sampled) { slots[i].n = n; slots[i].g0 = g0; full_slots.push_back(i); ++nbatch; }
                else free_slots.push_back(i);
                cv.notify_all();
            }
            if (n > 0) g0 += n;
            if (n < want) { short_file = true; break; }
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            done = true;
            cv.notify_all();
        }
        for (auto &w : workers) w.join();
        if (first_rc) { rc = first_rc; pa::set_error(first_err); }
        if (dbg_t) {
            fprintf(stderr, "[prealpha_b200] streaming: %lld batches of <= %d frames on %d device(s); reader %.3f s = read %.3f + wait for a free buffer %.3f\n",
                    nbatch, B, ndev, now_s() - t_loop0, t_read, t_wait);
            for (int d = 0; d < ndev; ++d)
                fprintf(stderr, "[prealpha_b200] streaming: device %d took %lld batches; session %.3f s, upload %.3f s, launch %.3f s\n",
                        ex.device + d, dealt[d], t_create[d], t_upload[d], t_run[d]);
        }
        if (std::find(sess.begin(), sess.end(), nullptr) != sess.end()) {                 // a session could not be created
            for (pa_session *s : sess) if (s) pa_session_destroy(s);
            return rc ? rc : PA_ERR_CUDA;
        }
        if (rc == 0) finish_cache_write(r, ts, src);              // (no-op unless a cache is being written)
        // sticky flags of every device, then the sum on the first one
        for (int d = 0; d < ndev && rc == 0; ++d) rc = pa_session_download(sess[d], nullptr, nullptr, nullptr, 0);
        if (comm_thread.joinable()) comm_thread.join();
        if (rc == 0) rc = pa_sessions_reduce_to_first(sess.data(), ndev);
        if (rc == 0) rc = pa_session_download(sess[0], hp.data(), within.data(), oob.data(), 0);
        if (st) pa_sessions_merge_stats(sess.data(), ndev, true, st);
        for (pa_session *s : sess) pa_session_destroy(s);
        if (rc == PA_ERR_UNSUPPORTED && !(ex.flags & PA_EXEC_NO_CELLS)) { ex.flags |= PA_EXEC_NO_CELLS; continue; }
        if (rc == PA_ERR_UNSUPPORTED && !(ex.flags & PA_EXEC_FORCE_GENERAL)) { ex.flags |= PA_EXEC_FORCE_GENERAL; continue; }
        if (rc == 0 && short_file) {                              // Molecular:5003-5009
            r.error(84, "trajectory is shorter than specified; using " + std::to_string(g0) + " steps");
            r.nsteps = g0;
        }
        return rc;
    }
    return PA_ERR_UNSUPPORTED;
}

int perform_distribution(Run &r, const std::string &input_name)     // Distribution:1187-1228
{
    if (r.nsteps < 11) { r.error(37, "trajectory is too short for the distribution analysis (< 11 steps)"); return 37; }
    DistInput d;
    int rc = read_distribution_input(r, input_name, d);
    if (rc) return rc;
    std::vector<pa_type> types; pa_traj traj;
    make_traj(r, types, traj);
    std::vector<pa_job> jobs(d.refs.size());
    for (size_t n = 0; n < d.refs.size(); ++n) {
        pa_job_request rq{};
        rq.mode = d.mode;
        for (int k = 0; k < 3; ++k) rq.ref_xyz[k] = d.refs[n][k];
        rq.ref_type = d.refs[n][3]; rq.obs_type = d.refs[n][4];
        rq.bin_a = d.bin_a; rq.bin_b = d.bin_b; rq.sampling_interval = d.sampling_interval;
        rq.weigh_charge = d.weigh_charge; rq.use_coc = d.use_coc; rq.subtract_uniform = d.subtract_uniform;
        rq.maxdist_optimize = d.maxdist_optimize; rq.maxdist = d.maxdist; rq.normalise_clm = d.normalise_clm;
        if (d.mode == PA_MODE_CHARGE_ARM) {                       // Distribution:1205-1207 (check_charges, Molecular:191-202)
            const HostType &h = r.types[rq.ref_type - 1];
            bool any = false;
            for (double q : h.atom_charge) if (std::fabs((float)q) > 0.001f) { any = true; break; }
            if (!any) r.error(127, "something is wrong with the atomic charges of this molecule type index");
        }
        rc = pa_job_setup(&traj, &rq, &jobs[n]);
        if (rc) { r.error(rc, pa_last_error()); return rc; }
    }
    if (jobs.empty()) return 0;
    const size_t nb = (size_t)jobs[0].bin_a * jobs[0].bin_b;
    std::vector<std::vector<int64_t>> hist(jobs.size(), std::vector<int64_t>(nb));
    std::vector<int64_t *> hp(jobs.size());
    for (size_t n = 0; n < jobs.size(); ++n) hp[n] = hist[n].data();
    std::vector<int64_t> within(jobs.size()), oob(jobs.size());
    pa_stats st{};
    const auto t0 = std::chrono::steady_clock::now();
    pa_exec ex = r.exec;
    if (ex.ndevices < 0) {
        // "all usable devices" is a default, not an order: a job below ~2e11 pair evaluations finishes on one GPU before
        // the other contexts and the communicators exist
        double evals = 0.0;
        for (const pa_job &jb : jobs) {
            if (jb.mode == PA_MODE_CHARGE_ARM) continue;
            double nobs = 0.0;
            for (size_t t = 0; t < r.types.size(); ++t) if (jb.obs_type < 1 || (int)t == jb.obs_type - 1) nobs += r.types[t].nmol;
            evals += (double)r.types[jb.ref_type - 1].nmol * nobs * ((r.nsteps + jb.sampling_interval - 1) / jb.sampling_interval);
        }
        if (evals < 2.0e11 || pa_device_count() - ex.device < 2) ex.ndevices = 1;
    }
    if (r.stream) rc = histogram_streaming(r, traj, jobs, hp, within, oob, &st, ex);
    else rc = pa_distribution_histogram_multi(&traj, jobs.data(), (int)jobs.size(), &ex, hp.data(), within.data(), oob.data(), &st);
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rc) { r.gpu_failed = true; r.error(rc, std::string("GPU histogram failed: ") + pa_last_error()); return rc; }
    traj.nsteps = r.nsteps;                                   // (a streamed file may have turned out shorter: error 84)
    printf("   ### B200 execution (distribution function): %lld frames, %lld pair evaluations, %lld kernel launches\n",
           (long long)st.frames_done, (long long)st.pairs_evaluated, (long long)st.kernel_launches);
    printf("   ### End of accelerated section, took %.3f seconds (kernels %.3f ms)\n", sec, st.kernel_ms);
    const std::string prefix = r.path_out + r.prefix;
    for (size_t n = 0; n < jobs.size(); ++n) {
        printf(" Calculating distribution for reference vector %zu out of %zu.\n", n + 1, jobs.size());
        printf("   within bin %lld, out of bounds %lld.\n", (long long)within[n], (long long)oob[n]);   // Distribution:911
        std::vector<float> g(nb);
        float rho = 0.f;
        rc = pa_transfer_histogram(&traj, &jobs[n], hist[n].data(), within[n], oob[n], r.verbose ? 1 : 0, g.data(), &rho);
        if (rc == PA_E_BIN_OVERFLOW) r.error(101, "integer overflow in bin count");
        if (r.verbose) printf("   ideal density:%10.3E particles/Angstrom^3\n", (double)rho);
        rc = pa_write_distribution(&traj, &jobs[n], (int)n + 1, prefix.c_str(), input_name.c_str(), g.data(), rho);
        if (rc) { r.error(rc, pa_last_error()); return rc; }
    }
    return 0;
}


// ---- `distance <file>`: Distances:299-802 (input, wildcard expansion), :1093-1159 (report) -------------
struct DistSubjob {
    std::vector<int32_t> ref, obs;          // (type, atom) pairs, 1-based
    long long weight = 0;
    std::string description, el1, el2;
};

std::vector<std::array<int, 2>> atoms_of_element(const Run &r, const std::string &el, int only_type /* 1-based or 0 */)
{
    std::vector<std::array<int, 2>> v;      // give_indices_of_specific_atoms[_per_molecule]: type-major, then atom order
    for (size_t t = 0; t < r.types.size(); ++t) {
        if (only_type && (int)t + 1 != only_type) continue;
        for (int a = 0; a < r.types[t].natoms; ++a)
            if (r.types[t].elements[a] == el) v.push_back({ (int)t + 1, a + 1 });
    }
    return v;
}

// Fortran Ew.d edit descriptor (0.ddddE+ee, right-justified in w): the digits are the correctly rounded decimal
// expansion of the binary value (exact ties to even, like libgfortran), taken from printf's %.{d-1}e
std::string fortran_E(double v, int w, int d)
{
    char buf[64];
    std::string body;
    if (std::isnan(v)) body = "NaN";
    else if (std::isinf(v)) body = v < 0 ? "-Infinity" : "Infinity";
    else if (v == 0.0) body = std::string(std::signbit(v) ? "-0." : "0.") + std::string((size_t)d, '0') + "E+00";
    else {
        snprintf(buf, sizeof buf, "%.*e", d - 1, std::fabs(v));          // D.DDDDe+XX
        std::string digits;
        int ex = 0;
        const char *e = strchr(buf, 'e');
        for (const char *p = buf; p < e; ++p) if (*p >= '0' && *p <= '9') digits += *p;
        ex = atoi(e + 1) + 1;
        snprintf(buf, sizeof buf, "%s0.%sE%+03d", v < 0 ? "-" : "", digits.c_str(), ex);
        body = buf;
    }
    if ((int)body.size() < w) body = std::string((size_t)w - body.size(), ' ') + body;
    return body;
}

std::string fortran_e93(double v)           // E9.3 = 0.dddE+ee; correctly rounded digits of the binary value, exact ties to even
{
    if (std::isnan(v)) return "      NaN";
    return fortran_E(v, 9, 3);
}

void formatted_distance_output(const std::string &what, double d)   // Distances:1145-1157
{
    printf("     %s: ", what.c_str());
    if (d > 1000.0) printf("%s\n", fortran_e93(d).c_str());
    else if (d > 1.0) printf("%.3f\n", d);
    else if (d > 0.01) printf("%5.3f\n", d);
    else printf("%s\n", fortran_e93(d).c_str());
}

int perform_distance(Run &r, const std::string &input_name)
{
    std::vector<std::string> lines;
    if (!read_lines(r.path_in + input_name, lines)) { r.error(134, "distance input file '" + r.path_in + input_name + "' not found"); return 134; }
    std::vector<std::vector<std::string>> toks;
    for (auto &l : lines) { auto t = tokens_of(l); if (!t.empty()) toks.push_back(t); }
    if (toks.empty() || toks[0].size() < 2) { r.error(135, "incorrect format in distance input"); return 135; }
    std::string mode = toks[0][0];
    if (mode == "intramolecular") mode = "intra";
    if (mode == "intermolecular") mode = "inter";
    int nsub = 0;
    if ((mode != "intra" && mode != "inter") || !parse_int(toks[0][1], nsub)) { r.error(135, "incorrect format in distance input"); return 135; }
    const bool inter = mode == "inter";
    const int ntypes = (int)r.types.size();
    printf(" reading %d user-specified distances...\n", nsub);
    std::vector<DistSubjob> subs;
    auto natoms_of = [&](int t) { return (t >= 1 && t <= ntypes) ? r.types[t - 1].natoms : 0; };
    size_t line = 1;
    for (int n = 0; n < nsub && line < toks.size(); ++n, ++line) {
        const auto &tk = toks[line];
        int a = 0, b = 0, c = 0, d = 0;
        DistSubjob sj;
        int wild = -1, t1 = 0, a1 = 0, t2 = 0, a2 = 0;
        std::string e1, e2;
        if (inter) {
            if (tk.size() >= 4 && parse_int(tk[0], a) && parse_int(tk[1], b) && parse_int(tk[2], c) && parse_int(tk[3], d)) {
                t1 = a; a1 = b; t2 = c; a2 = d;
                if (a1 < 1 || a1 > natoms_of(t1) || a2 < 1 || a2 > natoms_of(t2)) { r.error(111, "invalid atom index - this line is ignored"); continue; }
                if (t1 == t2 && r.types[t1 - 1].nmol == 1) { r.error(141, "intermolecular distance needs more than one molecule"); continue; }
                wild = 0;
            } else if (tk.size() >= 3 && parse_int(tk[1], b) && parse_int(tk[2], c)) {                 // El type atom -> swapped
                e1 = tk[0].substr(0, 2);
                if (c < 1 || c > natoms_of(b)) { r.error(natoms_of(b) ? 111 : 110, "invalid index - this line is ignored"); continue; }
                if (atoms_of_element(r, e1, 0).empty()) { r.error(136, "no atoms of element '" + e1 + "'"); continue; }
                r.notice(140, "Swapped reference and observed atoms (wildcard -> specific not allowed).");
                wild = 1; t1 = b; a1 = c; e2 = e1;
            } else if (tk.size() >= 3 && parse_int(tk[0], a) && parse_int(tk[1], b)) {                 // type atom El
                e2 = tk[2].substr(0, 2);
                if (b < 1 || b > natoms_of(a)) { r.error(natoms_of(a) ? 111 : 110, "invalid index - this line is ignored"); continue; }
                if (atoms_of_element(r, e2, 0).empty()) { r.error(136, "no atoms of element '" + e2 + "'"); continue; }
                wild = 1; t1 = a; a1 = b;
            } else if (tk.size() >= 2) {                                                                // El El
                e1 = tk[0].substr(0, 2); e2 = tk[1].substr(0, 2);
                if (atoms_of_element(r, e1, 0).empty() || atoms_of_element(r, e2, 0).empty()) { r.error(136, "no atoms of that element"); continue; }
                wild = 2;
            } else break;
        } else {
            if (tk.size() >= 3 && parse_int(tk[0], a) && parse_int(tk[1], b) && parse_int(tk[2], c)) {
                t1 = t2 = a; a1 = b; a2 = c;
                if (a1 < 1 || a1 > natoms_of(t1) || a2 < 1 || a2 > natoms_of(t1) || a1 == a2) { r.error(natoms_of(t1) ? 111 : 110, "invalid index - this line is ignored"); continue; }
                wild = 0;
            } else if (tk.size() >= 3 && parse_int(tk[0], a) && parse_int(tk[2], c)) {                 // type El atom -> swapped
                e1 = tk[1].substr(0, 2);
                if (c < 1 || c > natoms_of(a)) { r.error(natoms_of(a) ? 111 : 110, "invalid index - this line is ignored"); continue; }
                if (atoms_of_element(r, e1, a).empty()) { r.error(136, "no atoms of element '" + e1 + "' in that molecule"); continue; }
                r.notice(140, "Swapped reference and observed atoms (wildcard -> specific not allowed).");
                wild = 1; t1 = a; a1 = c; e2 = e1;
            } else if (tk.size() >= 3 && parse_int(tk[0], a) && parse_int(tk[1], b)) {                 // type atom El
                e2 = tk[2].substr(0, 2);
                if (b < 1 || b > natoms_of(a)) { r.error(natoms_of(a) ? 111 : 110, "invalid index - this line is ignored"); continue; }
                if (atoms_of_element(r, e2, a).empty()) { r.error(136, "no atoms of element '" + e2 + "' in that molecule"); continue; }
                wild = 1; t1 = a; a1 = b;
            } else if (tk.size() >= 2) {
                e1 = tk[0].substr(0, 2); e2 = tk[1].substr(0, 2);
                bool ok = false;
                for (int t = 1; t <= ntypes && !ok; ++t) ok = !atoms_of_element(r, e1, t).empty() && !atoms_of_element(r, e2, t).empty();
                if (!ok) { r.error(136, "no molecule contains both elements"); continue; }
                wild = 2;
            } else break;
        }
        // prepare_subjob, Distances:500-609
        char buf[512];
        if (wild == 0) {
            sj.ref = { t1, a1 }; sj.obs = { t2, a2 };
            snprintf(buf, sizeof buf, "Atom #%d in molecule type %d -> atom #%d in molecule type %d", a1, t1, a2, t2);
            sj.weight = r.types[t1 - 1].nmol;
        } else if (wild == 1) {
            sj.ref = { t1, a1 };
            auto v = atoms_of_element(r, e2, inter ? 0 : t1);
            for (auto &x : v) { sj.obs.push_back(x[0]); sj.obs.push_back(x[1]); }
            snprintf(buf, sizeof buf, "Atom #%d in molecule type %d -> %zu %s atoms.", a1, t1, v.size(), e2.c_str());
            sj.weight = r.types[t1 - 1].nmol;
        } else {
            auto v1 = atoms_of_element(r, e1, 0), v2 = atoms_of_element(r, e2, 0);
            for (auto &x : v1) { sj.ref.push_back(x[0]); sj.ref.push_back(x[1]); }
            for (auto &x : v2) { sj.obs.push_back(x[0]); sj.obs.push_back(x[1]); }
            snprintf(buf, sizeof buf, "%zu %s atoms -> %zu %s atoms.", v1.size(), e1.c_str(), v2.size(), e2.c_str());
            for (auto &x : v1)                                                      // count_wildcard_weights, Distances:810-845
                for (auto &y : v2) {
                    if (!inter) { if (x[0] == y[0] && x[1] != y[1]) { sj.weight += r.types[x[0] - 1].nmol; break; } }
                    else if (x[0] != y[0] || x[1] != y[1]) { sj.weight += r.types[x[0] - 1].nmol; break; }
                    else if (r.types[x[0] - 1].nmol > 1) { sj.weight += r.types[x[0] - 1].nmol; break; }
                }
        }
        sj.description = buf;
        sj.el1 = r.types[sj.ref[0] - 1].elements[sj.ref[1] - 1];
        sj.el2 = r.types[sj.obs[0] - 1].elements[sj.obs[1] - 1];
        subs.push_back(sj);
    }
    printf(" Successfully read %zu/%d custom distance jobs.\n", subs.size(), nsub);
    if (subs.empty()) { r.error(139, "no valid distance jobs"); return 139; }
    // keyword block, Distances:366-489; defaults :9-14 and set_defaults :288-296
    int nsteps = 1, sampling = 1;
    bool stdev = false;
    float maxdist = 10.0f, exponent = 1.0f;
    for (line = 1 + (size_t)nsub; line < toks.size(); ++line) {
        const auto &tk = toks[line];
        const std::string &k = tk[0];
        bool lb; int iv; float fv;
        if (k == "sampling_interval") { if (tk.size() > 1 && parse_int(tk[1], iv)) sampling = iv; else { r.error(100, "cannot read 'sampling_interval'"); sampling = 1; } }
        else if (k == "exponent") { if (tk.size() > 1 && parse_real4(tk[1], fv)) exponent = fv; else { r.error(100, "cannot read 'exponent'"); exponent = 1.0f; } }
        else if (k == "exponential_weight" || k == "weigh_exponential" || k == "calculate_exponential" || k == "average_exponential") {
            if (tk.size() > 1 && parse_logical(tk[1], lb)) r.dist_calc_exp = lb; else { r.error(100, "cannot read '" + k + "'"); r.dist_calc_exp = false; }
        } else if (k == "stdev" || k == "standard_deviation" || k == "standarddev") {
            if (tk.size() > 1 && parse_logical(tk[1], lb)) stdev = lb; else { r.error(100, "cannot read '" + k + "'"); stdev = false; }
        } else if (k == "ffc_weight" || k == "weigh_ffc" || k == "calculate_ffc" || k == "average_ffc" || k == "ffc" || k == "dhh" || k == "rhh") {
            if (tk.size() > 1 && parse_logical(tk[1], lb)) r.dist_calc_ffc = lb; else { r.error(100, "cannot read '" + k + "'"); r.dist_calc_ffc = false; }
        } else if (k == "nsteps") { if (tk.size() > 1 && parse_int(tk[1], iv)) nsteps = iv; else { r.error(100, "cannot read 'nsteps'"); nsteps = 1; } }
        else if (k == "maxdist" || k == "cutoff") { if (tk.size() > 1 && parse_real4(tk[1], fv)) maxdist = fv; else { r.error(100, "cannot read 'maxdist'"); maxdist = 10.0f; } }
        else if (k == "maxdist_optimize" || k == "cutoff_optimize") {
            double lim = r.box_hi[0] - r.box_lo[0];
            for (int c = 1; c < 3; ++c) lim = std::min(lim, r.box_hi[c] - r.box_lo[c]);
            maxdist = r.box_given ? (float)(lim / 2.0) : 10.0f;
        } else if (k == "quit") break;
    }
    if (nsteps > r.nsteps) nsteps = r.nsteps;
    if (nsteps < 1) nsteps = 1;
    if (sampling < 1) sampling = 1;
    const int nav = std::max((nsteps - 1 + sampling) / sampling, 0);
    printf(" Using %d timesteps in intervals of %d for averaging.\n calculating average %smolecular distances.\n", nav, sampling, mode.c_str());
    std::vector<pa_type> types; pa_traj traj;
    make_traj(r, types, traj);
    std::vector<pa_dist_subjob> csj(subs.size());
    for (size_t i = 0; i < subs.size(); ++i) {
        csj[i].n_ref = (int)subs[i].ref.size() / 2; csj[i].n_obs = (int)subs[i].obs.size() / 2;
        csj[i].ref = subs[i].ref.data(); csj[i].obs = subs[i].obs.data(); csj[i].weight = subs[i].weight;
    }
    pa_dist_job job{};
    job.intermolecular = inter ? 1 : 0; job.nsubjobs = (int)csj.size(); job.subjobs = csj.data();
    job.nsteps = nsteps; job.sampling_interval = sampling;
    job.calc_exp = r.dist_calc_exp; job.calc_ffc = r.dist_calc_ffc; job.calc_stdev = stdev;
    job.maxdist = maxdist; job.exponent = exponent;
    std::vector<pa_dist_result> res(csj.size());
    const int rc = pa_distance_subjobs(&traj, &job, &r.exec, res.data());
    if (rc) { r.gpu_failed = true; r.error(rc, std::string("GPU distance analysis failed: ") + pa_last_error()); return rc; }
    // report_distances, Distances:1093-1143
    printf(" Done with %smolecular distance calculation. Results:\n", mode.c_str());
    for (size_t i = 0; i < subs.size(); ++i) {
        printf("   Subjob #%zu out of %zu:\n     %s\n", i + 1, subs.size(), subs[i].description.c_str());
        if (res[i].missed_neighbours > 0) printf("     %lld atoms without neighbour - consider increasing cutoff.\n", (long long)res[i].missed_neighbours);
        else printf("     Used %lld reference atoms per step.\n", subs[i].weight);
        formatted_distance_output("average closest distance", res[i].closest_average);
        if (stdev) formatted_distance_output("standard deviation", res[i].closest_stdev);
        if (r.dist_calc_exp) {
            formatted_distance_output("exponentially weighed distance", res[i].exponential_average);
            if (stdev) formatted_distance_output("standard deviation", res[i].exponential_stdev);
        }
        if (r.dist_calc_ffc) {
            const std::string l = inter ? "d" : "r";
            formatted_distance_output("FFC distance " + l + subs[i].el1 + subs[i].el2, res[i].ffc_average);
            formatted_distance_output("exp(-kr) weighed FFC distance " + l + "'" + subs[i].el1 + subs[i].el2, res[i].ffcexp_average);
            formatted_distance_output("Number of averages for FFC", res[i].ffc_weight);
        }
    }
    printf(" ( reference atom(s) -> observed atom(s) )\n");
    return 0;
}

// ---- `cluster` keyword (SURVEY 8f-4 consumer): Module_Cluster's statistics path ------------------------------------
// The reference tests every pair of listed atoms with give_smallest_atom_distance_squared(...) < cutoff**2 in a serial
// double loop per time step (Cluster:1092-1100 GLOBAL, :1411-1426 PAIRS) and merges clusters as it goes.  Pairs it skips
// because both atoms already share a cluster would have been no-ops, so the outcome depends only on the qualifying
// pairs and the order in which they are met: pa_neighbour_pairs (GPU cell list, same REAL(4) predicate) returns
// exactly those pairs in loop order, and the merge bookkeeping below replays Cluster:1101-1180 on them.  Written:
// `<prefix>cluster_statistics.dat` (Cluster:1296-1330 / :1640-1660), byte-identical to the reference's; the cluster
// xyz dumps (print_step) and the intermittent autocorrelation function are outside this path and are not produced.
struct ClusterAtom { int type, mol, atom; float weight; };
struct ClusterPair { int a0, a1, b0, b1, type_a, atom_a, type_b, atom_b; };

int perform_cluster(Run &r, const std::string &input_name)
{
    std::vector<std::string> lines;
    if (!read_lines(r.path_in + input_name, lines)) { r.error(170, "cluster input file '" + r.path_in + input_name + "' not found"); return 170; }
    std::vector<std::vector<std::string>> toks;
    for (auto &l : lines) { auto t = tokens_of(l); if (!t.empty()) toks.push_back(t); }
    if (toks.empty()) { r.error(177, "cannot read the operation mode of the cluster analysis"); return 177; }
    const std::string m0 = toks[0][0];
    bool pairs_mode;
    if (m0 == "PAIRS" || m0 == "Pairs" || m0 == "pairs" || m0 == "PAIR" || m0 == "Pair" || m0 == "pair") pairs_mode = true;
    else if (m0 == "GLOBAL" || m0 == "Global" || m0 == "global") pairs_mode = false;
    else { r.error(177, "cannot read the operation mode of the cluster analysis"); return 177; }
    printf(pairs_mode ? " PAIRS operation mode:\n Uses a set of pairs of atoms with their corresponding cutoffs.\n"
                      : " GLOBAL operation mode:\n Uses a set of atoms and a global cutoff, applied to any combination of atoms.\n");
    const int nt = (int)r.types.size();
    auto valid_type = [&](int t) { return t >= 1 && t <= nt; };
    auto valid_atom = [&](int t, int a) { return valid_type(t) && a >= 1 && a <= r.types[t - 1].natoms; };
    int nsteps = 1, sampling = 1, tmax = 10000;
    bool print_statistics = false;
    std::vector<ClusterAtom> atoms;
    std::vector<ClusterPair> pairs;
    std::vector<float> cutoffs;
    float current_weight = 0.0f;
    auto add_instances = [&](int type, int atom) {             // one list entry per molecule of the type
        const int first = (int)atoms.size();
        for (int m = 1; m <= r.types[type - 1].nmol; ++m) atoms.push_back({ type, m, atom, current_weight });
        return std::make_pair(first, (int)atoms.size() - 1);
    };
    for (size_t n = 1; n < toks.size(); ++n) {                 // read_body, Cluster:509-891 (keywords of the statistics path)
        const auto &t = toks[n];
        const std::string &k = t[0];
        int i1, i2, i3, i4;
        float f1, f2;
        bool b1;
        if (k == "quit") break;
        if (k == "nsteps") { if (t.size() > 1 && parse_int(t[1], i1)) nsteps = i1; else r.error(170, "cannot read 'nsteps' (cluster input)"); }
        else if (k == "sampling_interval") { if (t.size() > 1 && parse_int(t[1], i1)) sampling = i1; else r.error(170, "cannot read 'sampling_interval' (cluster input)"); }
        else if (k == "tmax" || k == "t_max") { if (t.size() > 1 && parse_int(t[1], i1)) tmax = i1; else { r.error(24, "cannot read 'tmax'"); tmax = 10000; } }
        else if (k == "print_statistics" || k == "statistics") { if (t.size() > 1 && parse_logical(t[1], b1)) print_statistics = b1; else r.error(170, "cannot read 'print_statistics'"); }
        else if (k == "custom_weight" || k == "custom_weights" || k == "weight" || k == "weights" || k == "set_weight" || k == "set_weights") {
            if (t.size() > 1 && parse_real4(t[1], f1)) current_weight = f1; else r.error(170, "cannot read 'custom_weight'");
        } else if (k == "single_cutoff" || k == "cutoff") {
            if (pairs_mode) { r.error(178, "keyword not available in this operation mode"); continue; }
            if (!(t.size() > 1 && parse_real4(t[1], f1))) { r.error(170, "cannot read 'single_cutoff'"); continue; }
            if (f1 > 0.0f) cutoffs.push_back(f1); else r.error(175, "invalid cutoff");
        } else if (k == "scan_cutoff") {
            if (pairs_mode) { r.error(178, "keyword not available in this operation mode"); continue; }
            if (!(t.size() > 3 && parse_real4(t[1], f1) && parse_real4(t[2], f2) && parse_int(t[3], i1))) { r.error(170, "cannot read 'scan_cutoff'"); continue; }
            if (i1 > 0 || f1 < 0.0f || f1 > f2) {              // [sic], Cluster:542
                const float interval = (f2 - f1) / (float)(i1 - 1);
                for (int m = 0; m < i1; ++m) cutoffs.push_back(f1 + (float)m * interval);
            } else r.error(175, "invalid cutoff");
        } else if (k == "add_molecule_type" || k == "add_molecule_type_index" || k == "add_molecule") {
            if (pairs_mode) { r.error(178, "keyword not available in this operation mode"); continue; }
            if (!(t.size() > 1 && parse_int(t[1], i1))) { r.error(170, "cannot read 'add_molecule_type'"); continue; }
            if (valid_type(i1))
                for (int m = 1; m <= r.types[i1 - 1].nmol; ++m)
                    for (int a = 1; a <= r.types[i1 - 1].natoms; ++a) atoms.push_back({ i1, m, a, current_weight });
        } else if (k == "add_atom_index" || k == "add_atom") {
            if (pairs_mode) { r.error(178, "keyword not available in this operation mode"); continue; }
            if (!(t.size() > 2 && parse_int(t[1], i1) && parse_int(t[2], i2))) { r.error(170, "cannot read 'add_atom_index'"); continue; }
            if (valid_atom(i1, i2)) add_instances(i1, i2);
        } else if (k == "pair" || k == "Pair" || k == "add_pair") {
            if (!pairs_mode) { r.error(178, "keyword not available in this operation mode"); continue; }
            if (!(t.size() > 5 && parse_int(t[1], i1) && parse_int(t[2], i2) && parse_int(t[3], i3) && parse_int(t[4], i4) && parse_real4(t[5], f1))) continue;
            if (!(valid_atom(i1, i2) && valid_atom(i3, i4))) continue;
            ClusterPair cp{ -1, -1, -1, -1, i1, i2, i3, i4 };
            // an (atom index, molecule type) that an earlier pair already listed reuses that part of the atom list
            // (Cluster:652-726: A against old A, then old B; B against old A, then old B; finally B against the new A)
            for (const ClusterPair &q : pairs) {
                if (q.type_a == i1 && q.atom_a == i2) { cp.a0 = q.a0; cp.a1 = q.a1; break; }
                if (q.type_b == i1 && q.atom_b == i2) { cp.a0 = q.b0; cp.a1 = q.b1; break; }
            }
            for (const ClusterPair &q : pairs) {
                if (q.type_a == i3 && q.atom_a == i4) { cp.b0 = q.a0; cp.b1 = q.a1; break; }
                if (q.type_b == i3 && q.atom_b == i4) { cp.b0 = q.b0; cp.b1 = q.b1; break; }
            }
            if (cp.a0 < 0) { auto ab = add_instances(i1, i2); cp.a0 = ab.first; cp.a1 = ab.second; }
            if (cp.b0 < 0 && i3 == i1 && i4 == i2) { cp.b0 = cp.a0; cp.b1 = cp.a1; }
            if (cp.b0 < 0) { auto ab = add_instances(i3, i4); cp.b0 = ab.first; cp.b1 = ab.second; }
            pairs.push_back(cp);
            cutoffs.push_back(f1);
        }
        // print_step, print_spectators, autocorrelation, tmax, use_log: outputs outside this path (see above)
    }
    const int N_atoms = (int)atoms.size(), N_cutoffs = (int)cutoffs.size();
    if (N_atoms == 0) { r.error(173, "no atoms specified for the cluster analysis"); return 173; }
    if (N_cutoffs == 0) { r.error(174, "no cutoffs specified for the cluster analysis"); return 174; }
    if (nsteps < 1) { r.error(57, "timestep out of range"); nsteps = 1; }
    else if (nsteps > r.nsteps) { r.error(57, "timestep out of range"); nsteps = r.nsteps; }
    if (sampling < 1) sampling = 1;
    // the largest time shift of the (not built) autocorrelation function is checked whether or not it is computed (Cluster:887-890)
    if (tmax > std::max((nsteps - 1 + sampling) / sampling, 0) - 1 || tmax < 1) r.error(28, "tmax is too large (or too small)");
    if (pairs_mode) printf("   Expecting %d pairs and %d atoms.\n", (int)pairs.size(), N_atoms);
    else printf("   Expecting %d atoms and %d cutoffs.\n", N_atoms, N_cutoffs);
    if (print_statistics) {
        // custom_weights is never set by the reference's reader (Cluster:28,206,331): the molecular charge is always used
        printf(" no custom weights have been specified, using molecular charge.\n (Note: even for more than one atom, they are all assigned the full charge)\n");
        for (ClusterAtom &a : atoms) a.weight = (float)r.types[a.type - 1].charge;
    }
    if (!pairs_mode) std::sort(cutoffs.begin(), cutoffs.end());
    const int navg = std::max((nsteps - 1 + sampling) / sampling, 0);
    printf(" Using %d timesteps in intervals of %d for averaging.\n Starting cluster analysis.\n", navg, sampling);
    printf(" (B200 drop-in: neighbour pairs on the GPU, cluster bookkeeping replayed in the reference's order; the xyz dumps\n"
           "  of print_step and the intermittent autocorrelation function are not part of the accelerated path.)\n");
    std::vector<pa_type> types; pa_traj traj;
    make_traj(r, types, traj);
    const int ncol = pairs_mode ? 1 : N_cutoffs;
    std::vector<long long> occurrence((size_t)(N_atoms + 1) * ncol, 0), ccdf((size_t)ncol, 0);
    std::vector<double> avg_weight((size_t)(N_atoms + 1) * ncol, 0.0);
    std::vector<int32_t> inst((size_t)N_atoms * 3);
    for (int i = 0; i < N_atoms; ++i) { inst[3 * i] = atoms[i].type; inst[3 * i + 1] = atoms[i].mol; inst[3 * i + 2] = atoms[i].atom; }
    std::vector<int> cluster_number((size_t)N_atoms + 1);
    std::vector<float> cl_weight((size_t)N_atoms + 1);
    std::vector<int> cl_members((size_t)N_atoms + 1);
    std::vector<pa_neigh_pair> found((size_t)1 << 16);
    long long tests_total = 0, pairs_total = 0;
    const auto t0 = std::chrono::steady_clock::now();
    for (int step = 1; step <= nsteps; step += sampling) {
        int current = 0;
        std::fill(cluster_number.begin(), cluster_number.end(), 0);
        std::fill(cl_weight.begin(), cl_weight.end(), 0.0f);
        std::fill(cl_members.begin(), cl_members.end(), 0);
        auto merge = [&](int m, int n) {                      // Cluster:1101-1180 (1-based list positions)
            if (cluster_number[m] != 0 && cluster_number[m] == cluster_number[n]) return;
            if (cluster_number[m] == 0) {
                if (cluster_number[n] == 0) {
                    ++current;
                    cluster_number[m] = cluster_number[n] = current;
                    cl_weight[current] = atoms[m - 1].weight + atoms[n - 1].weight;
                    cl_members[current] = 2;
                } else {
                    cluster_number[m] = cluster_number[n];
                    cl_weight[cluster_number[n]] = cl_weight[cluster_number[n]] + atoms[m - 1].weight;
                    cl_members[cluster_number[n]] += 1;
                }
            } else if (cluster_number[n] == 0) {
                cluster_number[n] = cluster_number[m];
                cl_weight[cluster_number[m]] = cl_weight[cluster_number[m]] + atoms[n - 1].weight;
                cl_members[cluster_number[m]] += 1;
            } else {
                const int del = cluster_number[n], keep = cluster_number[m];
                cl_members[keep] += cl_members[del];
                cl_weight[keep] = cl_weight[keep] + cl_weight[del];
                for (int i = 1; i <= N_atoms; ++i) if (cluster_number[i] == del) cluster_number[i] = keep;
                if (current != del) {                          // the cluster with the highest number takes the free slot
                    cl_weight[del] = cl_weight[current];
                    cl_members[del] = cl_members[current];
                    for (int i = 1; i <= N_atoms; ++i) if (cluster_number[i] == current) cluster_number[i] = del;
                }
                --current;
            }
        };
        auto neighbour_pairs = [&](const int32_t *a, int na, const int32_t *b, int nb, bool same, float cutoff, int64_t &count) -> int {
            const float csq = cutoff * cutoff;                 // squared_cutoff_list(i)=cutoff_list(i)**2, REAL(4)
            pa_neigh_request rq{};
            rq.timestep = step; rq.n_a = na; rq.n_b = nb; rq.a = a; rq.b = b;
            rq.n_class_a = rq.n_class_b = 1; rq.cutoff_sq = &csq; rq.same_list = same ? 1 : 0; rq.skip_same_molecule = 0;
            for (;;) {
                int64_t tests = 0;
                const int rc = pa_neighbour_pairs(&traj, &rq, &r.exec, found.data(), (int64_t)found.size(), &count, &tests);
                if (rc) return rc;
                if (count <= (int64_t)found.size()) { tests_total += tests; pairs_total += count; return 0; }
                found.resize((size_t)count);
            }
        };
        auto statistics = [&](int col) {                       // Cluster:1185-1205 / :1516-1543
            long long counts = current;
            for (int i = 1; i <= current; ++i) {
                avg_weight[(size_t)cl_members[i] * ncol + col] += (double)cl_weight[i];
                occurrence[(size_t)cl_members[i] * ncol + col] += 1;
            }
            for (int i = 1; i <= N_atoms; ++i)
                if (cluster_number[i] == 0) {
                    ++counts;
                    occurrence[(size_t)1 * ncol + col] += 1;
                    avg_weight[(size_t)1 * ncol + col] += (double)atoms[i - 1].weight;
                }
            ccdf[col] += counts;
        };
        int rc = 0;
        int64_t count = 0;
        if (pairs_mode) {
            for (size_t p = 0; p < pairs.size() && rc == 0; ++p) {
                const ClusterPair &cp = pairs[p];
                rc = neighbour_pairs(inst.data() + 3 * (size_t)cp.a0, cp.a1 - cp.a0 + 1, inst.data() + 3 * (size_t)cp.b0, cp.b1 - cp.b0 + 1, false, cutoffs[p], count);
                for (int64_t q = 0; q < count && rc == 0; ++q) {
                    const int m = cp.a0 + found[q].ia + 1, n = cp.b0 + found[q].ib + 1;
                    if (m == n) continue;                      // the same atom (Cluster:1427-1429)
                    merge(m, n);
                }
            }
            if (rc == 0 && print_statistics) statistics(0);
        } else {
            for (int c = 0; c < N_cutoffs && rc == 0; ++c) {   // ascending cutoffs build on the clusters of the previous one
                rc = neighbour_pairs(inst.data(), N_atoms, inst.data(), N_atoms, true, cutoffs[c], count);
                for (int64_t q = 0; q < count && rc == 0; ++q) merge(found[q].ia + 1, found[q].ib + 1);
                if (rc == 0 && print_statistics) statistics(c);
            }
        }
        if (rc) { r.gpu_failed = true; r.error(rc, std::string("GPU neighbour search failed: ") + pa_last_error()); return rc; }
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("   ### B200 execution (cluster analysis): %d steps, %lld distance tests, %lld neighbour pairs, %.3f seconds\n", navg, tests_total, pairs_total, sec);
    if (print_statistics) {
        const std::string fn = r.path_out + r.prefix + "cluster_statistics.dat";
        if (r.verbose) printf(" Writing cluster statistics to '%s'\n", fn.c_str());
        FILE *f = fopen(fn.c_str(), "w");
        if (!f) { r.error(26, "cannot open '" + fn + "'"); return 26; }
        fprintf(f, " This file contains cluster statistics based on the input file '%s'\n", input_name.c_str());
        fprintf(f, " atom_fraction    = #(specified atoms in this cluster size) / total #(specified atoms)\n");
        fprintf(f, " cluster_fraction = #(clusters of this size) / total number of clusters\n");
        if (!pairs_mode) {
            if (N_cutoffs == 1) fprintf(f, " The cutoff was %s\n", fortran_E((double)cutoffs[0], 9, 3).c_str());
            else fprintf(f, "%9s", "cutoff");
        }
        fprintf(f, "%8s%15s%15s%17s\n", "members", "weight_average", "atom_fraction", "cluster_fraction");
        for (int c = 0; c < ncol; ++c) {
            float normalisation = 0.0f;
            long long sum_occ = 0;
            for (int i = 1; i <= N_atoms; ++i) {
                normalisation = normalisation + (float)(occurrence[(size_t)i * ncol + c] * i);
                sum_occ += occurrence[(size_t)i * ncol + c];
            }
            for (int i = 1; i <= N_atoms; ++i) {
                const long long occ = occurrence[(size_t)i * ncol + c];
                if (occ <= 0) continue;
                if (!pairs_mode && N_cutoffs > 1) fprintf(f, "%s", fortran_E((double)cutoffs[c], 9, 3).c_str());
                const double wavg = avg_weight[(size_t)i * ncol + c] / (double)(float)occ;
                const float afrac = (float)(occ * i) / normalisation;
                const float cfrac = (float)occ / (float)sum_occ;
                fprintf(f, "%8d%s%s%s\n", i, fortran_E(wavg, 15, 6).c_str(), fortran_E((double)afrac, 15, 6).c_str(), fortran_E((double)cfrac, 17, 6).c_str());
            }
        }
        fclose(f);
        if (pairs_mode) {
            const float ccf = (float)ccdf[0] / (float)navg;
            if (ccf < 0.01f || ccf > 99.99f) printf(" The 'CCF' for this analysis was%s\n", fortran_E((double)ccf, 11, 4).c_str());
            else printf(" The 'CCF' for this analysis was%6.2f\n", (double)ccf);
            printf(" ( CCF = cluster count function = average number of separate clusters per box)\n");
        } else {
            printf(" The following values of the CCF were observed:\n%10s%10s\n", "cutoff", "CCF");
            for (int c = 0; c < N_cutoffs; ++c) {
                const float ccf = (float)ccdf[c] / (float)navg;
                auto fmt = [](float v) { char b[32]; if (v < 0.01f || !(v < 99999.9f)) return fortran_E((double)v, 10, 4); snprintf(b, sizeof b, "%10.3f", (double)v); return std::string(b); };
                printf("%s%s\n", fmt(cutoffs[c]).c_str(), fmt(ccf).c_str());
            }
            printf(" ( CCF = cluster count function, for comparison with TRAVIS:\n J. Chem. Inf. Model., 2022, 62, 5634-5644, dx.doi.org/10.1021/acs.jcim.2c01244 )\n");
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------
// `speciation` keyword (Module_Speciation, statistics path): donor / acceptor coordination species.
//   input      read_acceptors_and_donors (Speciation:978-1256), read_body (:400-679: atom groups, N_species, N_neighbours,
//              nsteps, sampling_interval, use_ccc; the switches of the outputs that are not built are read and ignored)
//   neighbours the REAL(4) test of Speciation:1420-1428 for every acceptor atom x donor atom of a step, cutoffs
//              (r_acceptor + r_donor)**2 -- ONE pa_neighbour_pairs call per step and acceptor molecule type on the GPU
//   replay     the species clipboard of every acceptor molecule in the reference's loop order (:1383-1483), collapse to
//              atom groups and ordering (:1990-2078), recognition of an already known species by the greedy molecule
//              matching of :1778-1870, occurrence counts, communal cluster correction (:1509-1590, :1660-1674)
//   output     <prefix>acceptorN_speciation_statistics.dat and ..._speciation_statistics_table.dat (:2082-2885)
// Not built: the first / last structure dumps (speciesN_first.xyz ...), the species time series and its intermittent
// autocorrelation function, `group_elements` and covalence-radius cutoffs (negative radii).
struct SpecConn { int n_donor, dmol, dentry, aentry; };       // 1-based; dentry / aentry = position in the atom_index_list
struct SpecSide {
    int type = 0;                                              // molecule_type_index
    std::vector<int> atom, group;                              // atom_index_list(:,1), (:,2): own index, or -group number
    std::vector<float> radius;                                 // cutoff_list
    int n_groups = 0;
};
struct SpecMolRange { int first, extra; bool unmatched; };     // neighbour_molecule_starting_indices
struct SpecSpecies {
    int nconn = 0, natoms = 0, nmol = 0;
    long long occurrences = 0;
    std::vector<std::array<int, 4>> conn;                      // collapsed + ordered: donor type, donor molecule, donor group, acceptor group
    std::vector<SpecMolRange> mol;
    std::vector<std::array<double, 2>> ccc;                    // per donor type: sum of 1/sharing acceptors, connections
};
struct SpecDvs { int id, n_donor, parent; };

static std::string sum_formula(const HostType &t)             // give_sum_formula, Molecular:2194-2222
{
    std::string out;
    std::vector<char> used(t.elements.size(), 0);
    for (size_t o = 0; o < t.elements.size(); ++o) {
        if (used[o]) continue;
        int n = 1;
        for (size_t i = o + 1; i < t.elements.size(); ++i) if (t.elements[i] == t.elements[o]) { used[i] = 1; ++n; }
        out += t.elements[o];
        if (n > 1) out += std::to_string(n);
    }
    return out;
}

// source of the connections of one step and acceptor type: per acceptor molecule, in the reference's loop order
// (donor type, donor molecule, donor atom, acceptor atom)
using SpecSource = std::function<int(int step, int n_acceptor, std::vector<std::vector<SpecConn>> &per_molecule)>;

int perform_speciation(Run &r, const std::string &input_name)
{
    std::vector<std::string> lines;
    if (!read_lines(r.path_in + input_name, lines)) { r.error(158, "speciation input file '" + r.path_in + input_name + "' not found"); return 158; }
    std::vector<std::vector<std::string>> toks;
    for (auto &l : lines) { auto t = tokens_of(l); if (!t.empty()) toks.push_back(t); }
    const int nt = (int)r.types.size();
    size_t ln = 0;
    std::vector<SpecSide> acc, don;
    auto read_side = [&](std::vector<SpecSide> &side, const char *what) -> int {
        int n = 0;
        if (ln >= toks.size() || !parse_int(toks[ln][0], n)) { r.error(155, "incorrect format in speciation input"); return 155; }
        ++ln;
        if (r.verbose) printf(" reading %d user-specified %s atoms...\n", n, what);
        int good = 0;
        for (int k = 0; k < n; ++k, ++ln) {
            int ty, at; float cut;
            if (ln >= toks.size() || toks[ln].size() < 3 || !parse_int(toks[ln][0], ty) || !parse_int(toks[ln][1], at) || !parse_real4(toks[ln][2], cut)) {
                r.error(155, "incorrect format in speciation input"); return 155;
            }
            if (ty < 1 || ty > nt) { r.error(33, "invalid molecule type index in speciation input"); return 33; }
            if (at < 1 || at > r.types[ty - 1].natoms) { r.error(111, "invalid atom index in speciation input"); continue; }
            if (cut < -0.1f) { r.error(158, "covalence-radius cutoffs (negative radius) are not available in the B200 drop-in"); return 158; }
            SpecSide *s = nullptr;
            for (SpecSide &q : side) if (q.type == ty) s = &q;
            if (!s) { side.emplace_back(); s = &side.back(); s->type = ty; }
            s->atom.push_back(at); s->group.push_back(at); s->radius.push_back(cut);
            ++good;
        }
        if (good < 1) { r.error(159, std::string("no valid ") + what + " atoms"); return 159; }
        if (r.verbose) {
            printf(" Successfully read %d %s atoms:\n", good, what);
            for (const SpecSide &q : side) {
                printf("    Molecule type %d (%s):\n", q.type, sum_formula(r.types[q.type - 1]).c_str());
                for (size_t i = 0; i < q.atom.size(); ++i) {
                    printf("      Atom index %d (%s), cutoff ", q.atom[i], r.types[q.type - 1].elements[q.atom[i] - 1].c_str());
                    if (q.radius[i] < 0.01f) printf("< 0.01\n"); else printf("= %.2f\n", (double)q.radius[i]);
                }
            }
        }
        return 0;
    };
    if (int rc = read_side(acc, "acceptor")) return rc;
    if (int rc = read_side(don, "donor")) return rc;
    // ---- read_body --------------------------------------------------------------------------------------
    int nsteps = 100, sampling = 1, max_species = 11, max_conn = 10, group_type = 0, tmax = 10000;
    bool atoms_grouped = false, use_ccc = false;
    for (; ln < toks.size(); ++ln) {
        const auto &t = toks[ln];
        const std::string &k = t[0];
        int i1; bool b1;
        if (k == "quit") break;
        if (k == "maximum_number_of_connections" || k == "neighbours" || k == "N_neighbours") {
            if (t.size() > 1 && parse_int(t[1], i1)) max_conn = i1; else r.error(158, "cannot read 'maximum_number_of_connections'");
        } else if (k == "maximum_number_of_species" || k == "species" || k == "N_species") {
            if (t.size() > 1 && parse_int(t[1], i1)) max_species = i1; else r.error(158, "cannot read 'maximum_number_of_species'");
        } else if (k == "nsteps" || k == "n_steps") {
            if (t.size() > 1 && parse_int(t[1], i1)) nsteps = i1; else r.error(158, "cannot read 'nsteps'");
        } else if (k == "sampling_interval") {
            if (t.size() > 1 && parse_int(t[1], i1)) sampling = i1; else r.error(158, "cannot read 'sampling_interval'");
        } else if (k == "tmax") {
            if (t.size() > 1 && parse_int(t[1], i1)) tmax = i1; else { r.error(24, "cannot read 'tmax'"); tmax = 10000; }
        } else if (k == "ccc" || k == "use_ccc" || k == "cluster_correction" || k == "use_cluster_correction" || k == "communal_solvation" ||
                   k == "use_communal_solvation" || k == "use_communal_cluster_correction" || k == "communal_cluster_correction") {
            if (t.size() > 1 && parse_logical(t[1], b1)) use_ccc = b1; else r.error(158, "cannot read 'communal_cluster_correction'");
        } else if (k == "group_elements" || k == "group_element_names" || k == "group_element_symbols") {
            r.error(158, "'group_elements' is not available in the B200 drop-in"); return 158;
        } else if (k == "new_acceptor_group" || k == "acceptor_new_group" || k == "acceptor_group" || k == "acceptors_new_group" ||
                   k == "new_donor_group" || k == "donor_new_group" || k == "donor_group" || k == "donors_new_group") {
            const bool is_acc = k.find("acceptor") != std::string::npos;
            if (!(t.size() > 1 && parse_int(t[1], i1))) { r.error(158, "cannot read the molecule type of the new group"); continue; }
            group_type = i1;
            if (!atoms_grouped) {
                printf("   turned on atom groups.\n");
                for (SpecSide &q : (is_acc ? acc : don)) q.n_groups = 0;
            }
            bool ok = false;
            std::vector<SpecSide> &side = is_acc ? acc : don;
            for (size_t n = 0; n < side.size(); ++n)
                if (side[n].type == i1) {
                    ok = true;
                    side[n].n_groups += 1;
                    if (r.verbose) printf("   Added group %d to %s #%zu (molecule_type_index %d)\n", side[n].n_groups, is_acc ? "acceptor" : "donor", n + 1, i1);
                    atoms_grouped = true;
                }
            if (!ok) r.error(110, "no such molecule type among the " + std::string(is_acc ? "acceptors" : "donors"));
        } else if (k == "assign_to_acceptor_group" || k == "assign_atom_to_acceptor_group" || k == "add_to_acceptor_group" || k == "add_atom_to_acceptor_group" ||
                   k == "assign_to_donor_group" || k == "assign_atom_to_donor_group" || k == "add_to_donor_group" || k == "add_atom_to_donor_group") {
            const bool is_acc = k.find("acceptor") != std::string::npos;
            if (!atoms_grouped || group_type == 0) { r.error(164, "open a group before assigning atoms to it"); continue; }
            if (!(t.size() > 1 && parse_int(t[1], i1))) { r.error(158, "cannot read the atom index"); continue; }
            bool ok = false;
            for (SpecSide &q : (is_acc ? acc : don))
                if (q.type == group_type)
                    for (size_t i = 0; i < q.atom.size(); ++i)
                        if (q.atom[i] == i1) {
                            ok = true;
                            q.group[i] = -q.n_groups;
                            if (r.verbose) printf("   Added atom_index %d to %s group #%d (molecule_type_index %d)\n", i1, is_acc ? "acceptor" : "donor", q.n_groups, group_type);
                        }
            if (!ok) r.error(111, "invalid atom index in group assignment");
        }
        // print_beads, use_log, autocorrelation, tmax: switches of outputs that are not part of this path
    }
    if (nsteps < 1) { r.error(57, "timestep out of range"); nsteps = 1; }
    else if (nsteps > r.nsteps) { r.error(57, "timestep out of range"); nsteps = r.nsteps; }
    if (sampling < 1) sampling = 1;
    // the largest time shift of the (not built) autocorrelation function is checked whether or not it is computed (:673-677)
    if (tmax > std::max((nsteps - 1 + sampling) / sampling, 0) - 1 || tmax < 1) r.error(28, "tmax is too large (or too small)");
    if (max_conn < 1 || max_species < 1) { r.error(158, "N_species and N_neighbours must be positive"); return 158; }
    const int NA = (int)acc.size(), ND = (int)don.size();
    const int navg = std::max((nsteps - 1 + sampling) / sampling, 0);
    printf(" Using %d timesteps in intervals of %d for averaging.\n Starting speciation analysis.\n", navg, sampling);
    printf(" (B200 drop-in: neighbour pairs on the GPU, species bookkeeping replayed in the reference's order; structure dumps and\n"
           "  the species autocorrelation function are not part of the accelerated path.)\n");
    // ---- the neighbour stage ---------------------------------------------------------------------------------
    std::vector<int> acc_off(NA + 1, 0), don_off(ND + 1, 0);   // global entry index (cutoff classes)
    for (int n = 0; n < NA; ++n) acc_off[n + 1] = acc_off[n] + (int)acc[n].atom.size();
    for (int n = 0; n < ND; ++n) don_off[n + 1] = don_off[n] + (int)don[n].atom.size();
    std::vector<float> cut_sq((size_t)acc_off[NA] * don_off[ND]);
    for (int n = 0; n < NA; ++n)
        for (size_t a = 0; a < acc[n].atom.size(); ++a)
            for (int m = 0; m < ND; ++m)
                for (size_t b = 0; b < don[m].atom.size(); ++b) {
                    const float d = acc[n].radius[a] + don[m].radius[b];           // REAL :: real_space_distance, Speciation:386-389
                    cut_sq[(size_t)(acc_off[n] + (int)a) * don_off[ND] + (size_t)(don_off[m] + (int)b)] = d * d;
                }
    std::vector<int32_t> b_inst, b_cls;
    std::vector<int> b_ndonor, b_mol, b_entry;
    for (int m = 0; m < ND; ++m)
        for (int mol = 1; mol <= r.types[don[m].type - 1].nmol; ++mol)
            for (size_t b = 0; b < don[m].atom.size(); ++b) {
                b_inst.push_back(don[m].type); b_inst.push_back(mol); b_inst.push_back(don[m].atom[b]);
                b_cls.push_back(don_off[m] + (int)b);
                b_ndonor.push_back(m + 1); b_mol.push_back(mol); b_entry.push_back((int)b + 1);
            }
    std::vector<pa_type> types; pa_traj traj;
    make_traj(r, types, traj);
    std::vector<pa_neigh_pair> found((size_t)1 << 16);
    long long tests_total = 0, pairs_total = 0;
    SpecSource gpu_source = [&](int step, int n_acc, std::vector<std::vector<SpecConn>> &per_mol) -> int {
        const SpecSide &A = acc[n_acc - 1];
        const int nmol = r.types[A.type - 1].nmol, na_e = (int)A.atom.size();
        std::vector<int32_t> a_inst, a_cls;
        for (int mol = 1; mol <= nmol; ++mol)
            for (int a = 0; a < na_e; ++a) { a_inst.push_back(A.type); a_inst.push_back(mol); a_inst.push_back(A.atom[a]); a_cls.push_back(acc_off[n_acc - 1] + a); }
        pa_neigh_request rq{};
        rq.timestep = step; rq.n_a = nmol * na_e; rq.n_b = (int)b_cls.size(); rq.a = a_inst.data(); rq.b = b_inst.data();
        rq.a_class = a_cls.data(); rq.b_class = b_cls.data(); rq.n_class_a = acc_off[NA]; rq.n_class_b = don_off[ND];
        rq.cutoff_sq = cut_sq.data(); rq.same_list = 0; rq.skip_same_molecule = 1;     // Speciation:1394-1395
        int64_t count = 0;
        for (;;) {
            int64_t tests = 0;
            const int rc = pa_neighbour_pairs(&traj, &rq, &r.exec, found.data(), (int64_t)found.size(), &count, &tests);
            if (rc) return rc;
            if (count <= (int64_t)found.size()) { tests_total += tests; pairs_total += count; break; }
            found.resize((size_t)count);
        }
        per_mol.assign((size_t)nmol, {});
        for (int64_t q = 0; q < count; ++q) {
            const int ia = found[q].ia, ib = found[q].ib;
            per_mol[(size_t)(ia / na_e)].push_back({ b_ndonor[ib], b_mol[ib], b_entry[ib], ia % na_e + 1 });
        }
        // pairs arrive ordered by (acceptor atom, donor instance); the reference's loops run donor type, donor molecule,
        // donor atom, acceptor atom
        for (auto &v : per_mol)
            std::stable_sort(v.begin(), v.end(), [](const SpecConn &x, const SpecConn &y) {
                if (x.n_donor != y.n_donor) return x.n_donor < y.n_donor;
                if (x.dmol != y.dmol) return x.dmol < y.dmol;
                if (x.dentry != y.dentry) return x.dentry < y.dentry;
                return x.aentry < y.aentry;
            });
        return 0;
    };
    // test seam (pa_debug_run_general_input_with_connections): the connections come from a table the caller computed, so
    // that the bookkeeping below can be checked on a machine without a GPU; pa_run_general_input never takes this branch
    const int call_ordinal = r.spec_calls++;
    SpecSource table_source = [&](int step, int n_acc, std::vector<std::vector<SpecConn>> &per_mol) -> int {
        per_mol.assign((size_t)r.types[acc[n_acc - 1].type - 1].nmol, {});
        for (int64_t q = 0; q < r.dbg_nconns; ++q) {
            const int32_t *c = r.dbg_conns + 8 * q;
            if (c[0] != call_ordinal || c[1] != step || c[2] != n_acc) continue;
            if (c[3] < 1 || c[3] > (int)per_mol.size()) return PA_ERR_BAD_ARG;
            per_mol[(size_t)c[3] - 1].push_back({ c[4], c[5], c[6], c[7] });
        }
        for (auto &v : per_mol)
            std::stable_sort(v.begin(), v.end(), [](const SpecConn &x, const SpecConn &y) {
                if (x.n_donor != y.n_donor) return x.n_donor < y.n_donor;
                if (x.dmol != y.dmol) return x.dmol < y.dmol;
                if (x.dentry != y.dentry) return x.dentry < y.dentry;
                return x.aentry < y.aentry;
            });
        return 0;
    };
    const SpecSource &source = r.dbg_conns ? table_source : gpu_source;
    // ---- replay ------------------------------------------------------------------------------------------------
    std::vector<std::vector<SpecSpecies>> species(NA);
    std::vector<long long> none(NA, 0), species_overflow(NA, 0);
    long long neighbour_atom_overflow = 0;
    int maxmol = 0;
    for (const SpecSide &A : acc) maxmol = std::max(maxmol, r.types[A.type - 1].nmol);
    std::vector<SpecDvs> dvs((size_t)maxmol * max_conn + 2, SpecDvs{ 0, 0, 0 });     // donor_vs_species_list (persistent, :1298)
    std::vector<int> donor_id_off(ND + 1, 0);
    for (int m = 0; m < ND; ++m) donor_id_off[m + 1] = donor_id_off[m] + r.types[don[m].type - 1].nmol;
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::vector<SpecConn>> per_mol;
    for (int step = 1; step <= nsteps; step += sampling) {
        for (int n_acc = 1; n_acc <= NA; ++n_acc) {
            const SpecSide &A = acc[n_acc - 1];
            std::vector<SpecSpecies> &list = species[n_acc - 1];
            if (int rc = source(step, n_acc, per_mol)) { r.gpu_failed = true; r.error(rc, std::string("GPU neighbour search failed: ") + pa_last_error()); return rc; }
            int logged = 0, new_conn = 0;
            for (size_t am = 0; am < per_mol.size(); ++am) {
                // ---- clipboard of this acceptor molecule (:1383-1470) ----
                SpecSpecies clip;
                new_conn = 0;
                {
                    int cur_nd = -1, cur_mol = -1, cur_atom = -1;
                    bool atom_counted = false, mol_counted = false;
                    for (const SpecConn &c : per_mol[am]) {
                        if (c.n_donor != cur_nd || c.dmol != cur_mol) { cur_nd = c.n_donor; cur_mol = c.dmol; cur_atom = -1; mol_counted = false; }
                        if (c.dentry != cur_atom) { cur_atom = c.dentry; atom_counted = false; }
                        if (clip.nconn == max_conn) { ++neighbour_atom_overflow; continue; }
                        clip.conn.push_back({ c.n_donor, c.dmol, c.dentry, c.aentry });
                        ++clip.nconn;
                        if (!atom_counted) { atom_counted = true; ++clip.natoms; if (!mol_counted) { mol_counted = true; ++clip.nmol; } }
                        if (use_ccc) {
                            ++new_conn;
                            SpecDvs &d = dvs[(size_t)(logged + new_conn)];
                            d.id = donor_id_off[c.n_donor - 1] + c.dmol; d.n_donor = c.n_donor;
                        }
                    }
                }
                int observed = 0;
                if (clip.nconn == 0) { ++none[n_acc - 1]; }
                else {
                    // ---- collapse_and_sort_atom_indices (:1990-2078) ----
                    for (auto &c : clip.conn) { c[2] = don[c[0] - 1].group[c[2] - 1]; c[3] = A.group[c[3] - 1]; }
                    auto sort_level = [&](int level, int first, int last) {           // insertion sort, positions first..last (0-based, inclusive)
                        for (int i = first + 1; i <= last; ++i) {
                            const std::array<int, 4> x = clip.conn[i];
                            int j = i;
                            while (j > first && clip.conn[j - 1][level] > x[level]) { clip.conn[j] = clip.conn[j - 1]; --j; }
                            clip.conn[j] = x;
                        }
                    };
                    auto finish_molecule = [&](int first, int last) {
                        clip.mol.push_back({ first, last - first, false });
                        sort_level(2, first, last);
                        int afirst = first, cur = clip.conn[first][2];
                        for (int i = first; i < last; ++i)
                            if (clip.conn[i + 1][2] != cur) { sort_level(3, afirst, i); afirst = i + 1; cur = clip.conn[i + 1][2]; }
                        sort_level(3, afirst, last);
                    };
                    {
                        int first = 0;
                        for (int i = 0; i + 1 < clip.nconn; ++i)
                            if (clip.conn[i + 1][0] != clip.conn[first][0] || clip.conn[i + 1][1] != clip.conn[first][1]) { finish_molecule(first, i); first = i + 1; }
                        finish_molecule(first, clip.nconn - 1);
                    }
                    // ---- is it a known species?  (:1739-1772) ----
                    std::vector<char> potential(list.size(), 1);
                    for (size_t s = 0; s < list.size(); ++s)                          // eliminate_number_mismatches
                        if (list[s].nconn != clip.nconn || list[s].natoms != clip.natoms || list[s].nmol != clip.nmol) potential[s] = 0;
                    for (size_t s = 0; s < list.size(); ++s) {                        // eliminate_connection_mismatches
                        if (!potential[s]) continue;
                        SpecSpecies &ref = list[s];
                        bool success = false;
                        for (auto &m : clip.mol) m.unmatched = true;
                        for (auto &m : ref.mol) m.unmatched = true;
                        const int nm = clip.nmol;
                        for (int cm = 0; cm < nm && !success; ++cm) {
                            if (cm >= (int)clip.mol.size() || !clip.mol[cm].unmatched) continue;
                            for (int rm = 0; rm < nm; ++rm) {
                                if (rm >= (int)ref.mol.size()) break;
                                if (!ref.mol[rm].unmatched || !clip.mol[cm].unmatched) continue;
                                bool mismatch = false;
                                if (clip.mol[cm].extra != ref.mol[rm].extra) mismatch = true;
                                else
                                    for (int k = 0; k <= ref.mol[rm].extra; ++k) {
                                        const auto &x = ref.conn[(size_t)(ref.mol[rm].first + k)];
                                        const auto &y = clip.conn[(size_t)(clip.mol[cm].first + k)];
                                        if (x[0] != y[0] || x[2] != y[2] || x[3] != y[3]) { mismatch = true; break; }
                                    }
                                if (!mismatch) {
                                    clip.mol[cm].unmatched = false; ref.mol[rm].unmatched = false;
                                    bool any = false;
                                    for (auto &m : clip.mol) any = any || m.unmatched;
                                    for (auto &m : ref.mol) any = any || m.unmatched;
                                    if (!any) { success = true; break; }
                                }
                            }
                        }
                        potential[s] = success ? 1 : 0;
                    }
                    int match = -1;
                    for (size_t s = 0; s < list.size(); ++s) if (potential[s]) { match = (int)s; break; }
                    if (match < 0) {                                                  // append_species
                        if ((int)list.size() < max_species) {
                            clip.occurrences = 1;
                            clip.ccc.assign((size_t)ND, { 0.0, 0.0 });
                            list.push_back(clip);
                            observed = (int)list.size();
                        } else { ++species_overflow[n_acc - 1]; observed = 0; }
                    } else { list[(size_t)match].occurrences += 1; observed = match + 1; }
                }
                if (use_ccc && new_conn > 0 && observed > 0) {
                    for (int k = 1; k <= new_conn; ++k) dvs[(size_t)(logged + k)].parent = observed;
                    logged += new_conn;
                }
            }
            // ---- communal cluster correction of this step (:1509-1590) ----
            if (use_ccc && logged > 0) {
                if (logged == 1) {
                    // [sic] the donor type is read at position logged_connections + new_connections, new_connections being
                    // that of the LAST acceptor molecule of the loop (Speciation:1516-1523)
                    const SpecDvs &d = dvs[(size_t)std::min<long long>((long long)logged + new_conn, (long long)dvs.size() - 1)];
                    const int sp = dvs[1].parent;
                    if (sp >= 1 && sp <= (int)list.size() && d.n_donor >= 1 && d.n_donor <= ND) {
                        list[(size_t)sp - 1].ccc[(size_t)d.n_donor - 1][0] += 1.0;
                        list[(size_t)sp - 1].ccc[(size_t)d.n_donor - 1][1] += 1.0;
                    }
                } else {
                    std::vector<double> frac(list.size() * (size_t)ND, 0.0);
                    std::vector<int> cnt(list.size() * (size_t)ND, 0);
                    std::stable_sort(dvs.begin() + 1, dvs.begin() + 1 + logged, [](const SpecDvs &x, const SpecDvs &y) { return x.id < y.id; });
                    int first = 1;
                    auto chunk = [&](int f, int l) {
                        for (int n = f; n <= l; ++n) {
                            const size_t cell = (size_t)(dvs[(size_t)n].parent - 1) * ND + (size_t)(dvs[(size_t)n].n_donor - 1);
                            frac[cell] += (double)(1.0f / (float)(l - f + 1));                // 1.0/FLOAT(last-first+1): REAL(4)
                            cnt[cell] += 1;
                        }
                    };
                    for (int m = 2; m <= logged; ++m)
                        if (dvs[(size_t)m].id != dvs[(size_t)first].id) { chunk(first, m - 1); first = m; }
                    chunk(first, logged);
                    for (size_t s = 0; s < list.size(); ++s)
                        for (int n = 0; n < ND; ++n)
                            if (cnt[s * ND + n] != 0) { list[s].ccc[(size_t)n][0] += frac[s * ND + n]; list[s].ccc[(size_t)n][1] += (double)cnt[s * ND + n]; }
                }
            }
        }
    }
    if (use_ccc)
        for (auto &list : species) for (SpecSpecies &s : list) for (auto &c : s.ccc) c[0] = c[0] / c[1];          // (:1662-1672; 0/0 for absent donors)
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!r.dbg_conns)
        printf("   ### B200 execution (speciation): %d steps, %lld distance tests, %lld connections, %.3f seconds\n", navg, tests_total, pairs_total, sec);
    if (neighbour_atom_overflow > 0) r.error(160, "neighbour atom overflow (" + std::to_string(neighbour_atom_overflow) + "): increase N_neighbours");
    {
        long long so = 0;
        for (long long v : species_overflow) so += v;
        if (so > 0) r.notice(161, "species overflow (" + std::to_string(so) + "): increase N_species");
    }
    // ---- print_acceptor_summary (:2082-2885) ---------------------------------------------------------------------
    auto elem2 = [&](int type, int atom) {                       // CHARACTER(LEN=2) element symbol, untrimmed
        std::string e = r.types[type - 1].elements[atom - 1];
        e.resize(2, ' ');
        return e;
    };
    auto elem = [&](int type, int atom) { return r.types[type - 1].elements[atom - 1]; };
    auto formula = [&](int type) { return sum_formula(r.types[type - 1]); };
    auto describe_groups = [&](FILE *f, const SpecSide &S, const char *who) {
        for (int g = 1; g <= S.n_groups; ++g) {
            bool firsttime = true;
            for (size_t i = 0; i < S.atom.size(); ++i)
                if (g == -S.group[i]) {
                    if (firsttime) { firsttime = false; fprintf(f, "     Group #%d, atom indices %d(%s)", g, S.atom[i], elem(S.type, S.atom[i]).c_str()); }
                    else fprintf(f, ", %d(%s)", S.atom[i], elem(S.type, S.atom[i]).c_str());
                }
            if (firsttime) fprintf(f, "     Group #%d was empty.\n", g); else fprintf(f, ".\n");
        }
        bool none_ungrouped = true, firsttime = true;
        for (size_t i = 0; i < S.atom.size(); ++i)
            if (S.group[i] > 0) {
                none_ungrouped = false;
                if (firsttime) { fprintf(f, "     Atoms not assigned to a group, treated as their own group:\n       Atom indices %d", S.atom[i]); firsttime = false; }
                else fprintf(f, ", %d", S.atom[i]);
            }
        if (none_ungrouped) fprintf(f, "     All atoms of this %s were assigned to a group.\n", who); else fprintf(f, ".\n");
    };
    if (NA == 1) printf(" There was one acceptor: "); else printf(" There were %d acceptor molecules:\n", NA);
    for (int n_acc = 1; n_acc <= NA; ++n_acc) {
        const SpecSide &A = acc[n_acc - 1];
        std::vector<SpecSpecies> &list = species[n_acc - 1];
        if (NA > 1) printf("  Acceptor %d/%d was ", n_acc, NA);
        const std::string fn = r.path_out + r.prefix + "acceptor" + std::to_string(n_acc) + "_speciation_statistics.dat";
        FILE *f = fopen(fn.c_str(), "w");
        if (!f) { r.error(26, "cannot open '" + fn + "'"); return 26; }
        fprintf(f, " This file contains speciation statistics for acceptors coordinated by donors.\n");
        if (A.atom.size() == 1) fprintf(f, " This acceptor was atom %d (%s) of molecule type %d (%s).\n", A.atom[0], elem2(A.type, A.atom[0]).c_str(), A.type, formula(A.type).c_str());
        else if (atoms_grouped) {
            fprintf(f, " Molecule type %d (%s) was the acceptor.\n", A.type, formula(A.type).c_str());
            if (A.n_groups == 0) fprintf(f, "   This acceptor (#%d) has no custom groups defined.\n", n_acc);
            else fprintf(f, "   Groups for this acceptor (#%d):\n", n_acc);
            describe_groups(f, A, "acceptor");
        } else {
            fprintf(f, " Molecule type %d (%s) was the acceptor, with the following atom indices:\n", A.type, formula(A.type).c_str());
            for (size_t i = 0; i < A.atom.size(); ++i) fprintf(f, "      Atom index %d (%s)\n", A.atom[i], elem(A.type, A.atom[i]).c_str());
        }
        fprintf(f, " The donor molecule types and atom types are given at the bottom.\n");
        fprintf(f, " For this acceptor, %zu different species were observed,\n", list.size());
        fprintf(f, " Which are now given in order of decreasing probability:\n");
        printf("molecule type %d (%s).\n   For this acceptor, %zu different species were observed:\n", A.type, formula(A.type).c_str(), list.size());
        std::vector<float> occ(list.size());
        float total = 0.0f;
        for (size_t s = 0; s < list.size(); ++s) { occ[s] = (float)list[s].occurrences; total = total + occ[s]; }
        total = total + (float)species_overflow[n_acc - 1];
        total = total + (float)none[n_acc - 1];
        FILE *ft = nullptr;
        std::vector<int> formal((size_t)ND, 0);
        if (use_ccc) {
            const std::string fnt = r.path_out + r.prefix + "acceptor" + std::to_string(n_acc) + "_speciation_statistics_table.dat";
            ft = fopen(fnt.c_str(), "w");
            if (!ft) { fclose(f); r.error(26, "cannot open '" + fnt + "'"); return 26; }
            fprintf(ft, " This file contains the detailed communal cluster correction factors.\n");
            fprintf(ft, "%15s%15s", "#species", "<h>");
            for (int n = 0; n < ND; ++n) fprintf(ft, "%15s", formula(don[n].type).c_str());
            fprintf(ft, " sum_formula(formally)\n");
            if (none[n_acc - 1] > 0) {
                fprintf(ft, "%15d%15.10f", 0, (double)(species_overflow[n_acc - 1] + none[n_acc - 1]) / (double)total);
                for (int n = 0; n < ND; ++n) fprintf(ft, "%s", fortran_E(0.0, 15, 6).c_str());
                fprintf(ft, " [%s]\n", formula(A.type).c_str());
            }
        }
        int out_counter = 0;
        for (;;) {
            int best = -1; float best_value = 0.0f;
            for (size_t s = 0; s < list.size(); ++s) if (occ[s] > best_value) { best_value = occ[s]; best = (int)s; }
            if (best < 0) break;
            ++out_counter;
            occ[(size_t)best] = -1.0f;
            SpecSpecies &S = list[(size_t)best];
            const float share = (float)S.occurrences / total;
            fprintf(f, "   -------\n");
            if (share > 0.01f) fprintf(f, "   Species #(%d) in %.1f%% of the cases.\n", out_counter, (double)(100.0f * (float)S.occurrences / total));
            else fprintf(f, "     Species #(%d) in %lld cases (less than 1%%).\n", out_counter, S.occurrences);
            if (ft) { std::fill(formal.begin(), formal.end(), 0); fprintf(ft, "%15d%15.10f", out_counter, (double)share); }
            if (out_counter < 6) {
                if (share > 0.01f) printf("     Species #(%d) in %.1f%% of the cases.\n", out_counter, (double)(100.0f * (float)S.occurrences / total));
                else printf("     Species #(%d) in %lld cases (less than 1%%).\n", out_counter, S.occurrences);
            } else if (out_counter == 6) printf("     (Only the first 5 printed here)\n");
            fprintf(f, "   This species has a total of %d connections,\n", S.nconn);
            fprintf(f, "   with a total of %d neighbour atoms in %d neighbour molecules:\n", S.natoms, S.nmol);
            int charge = r.types[A.type - 1].charge;
            for (auto &m : S.mol) m.unmatched = true;
            for (int sm = 0; sm < S.nmol && sm < (int)S.mol.size(); ++sm) {
                const int first = S.mol[sm].first, extra = S.mol[sm].extra;
                charge += r.types[don[S.conn[(size_t)first][0] - 1].type - 1].charge;
                if (!S.mol[sm].unmatched) continue;
                int doubles = 1;
                for (int s2 = sm + 1; s2 < S.nmol && s2 < (int)S.mol.size(); ++s2) {
                    if (S.mol[s2].extra != extra) continue;
                    bool same = true;
                    for (int k = 0; k <= extra; ++k) {
                        const auto &x = S.conn[(size_t)(first + k)], &y = S.conn[(size_t)(S.mol[s2].first + k)];
                        if (x[2] != y[2] || x[3] != y[3] || x[0] != y[0]) { same = false; break; }
                    }
                    if (!same) continue;
                    ++doubles;
                    S.mol[s2].unmatched = false;
                }
                if (doubles > 1) fprintf(f, "     %d Molecules of type ", doubles); else fprintf(f, "     One molecule of type ");
                const int nd = S.conn[(size_t)first][0];
                if (ft) formal[(size_t)nd - 1] += doubles;
                fprintf(f, "%d(%s) with ", don[nd - 1].type, formula(don[nd - 1].type).c_str());
                if (extra == 0) fprintf(f, "one connection"); else fprintf(f, "%d connections", extra + 1);
                fprintf(f, doubles > 1 ? " each:\n" : ":\n");
                for (int k = 0; k <= extra; ++k) {
                    const int ag = S.conn[(size_t)(first + k)][3], dg = S.conn[(size_t)(first + k)][2];
                    if (dg < 0) fprintf(f, "       atom group %d (donor) ", -dg); else fprintf(f, "       atom index %d (donor) ", dg);
                    if (ag < 0) fprintf(f, "to atom group %d (acceptor)\n", -ag); else fprintf(f, "to atom index %d (acceptor)\n", ag);
                }
            }
            fprintf(f, "   The total formal charge including the acceptor was %+d.\n", charge);
            if (ft) {
                fprintf(f, "   The correction factors for communal solvation were:\n");
                for (int n = 0; n < ND; ++n) {
                    int per_donor = 0;
                    for (int k = 0; k < S.nconn; ++k) if (S.conn[(size_t)k][0] == n + 1) ++per_donor;
                    if (formal[(size_t)n] == 0) S.ccc[(size_t)n][0] = 0.0;
                    const double factor = S.ccc[(size_t)n][0], corr = factor * (double)(float)per_donor;
                    fprintf(ft, "%s", fortran_E(corr, 15, 6).c_str());
                    if (per_donor > 0) {
                        char b1[32], b2[32];
                        snprintf(b1, sizeof b1, "%4.2f", factor);
                        if (corr < 1.0) snprintf(b2, sizeof b2, "%4.2f", corr); else snprintf(b2, sizeof b2, "%.2f", corr);
                        fprintf(f, "     x%s for molecule type %d corresponding to %s %s\n", b1, n + 1, b2, formula(don[n].type).c_str());
                    }
                }
                fprintf(ft, " [%s][", formula(A.type).c_str());
                for (int n = 0; n < ND; ++n) fprintf(ft, "(%s)%d", formula(don[n].type).c_str(), formal[(size_t)n]);
                fprintf(ft, "]\n");
            }
        }
        if (species_overflow[n_acc - 1] > 0) {
            const float so = (float)species_overflow[n_acc - 1] / total;
            if (so > 0.01f) printf("     There was species overflow amounting to %.1f%%\n", (double)(100.0f * (float)species_overflow[n_acc - 1] / total));
            else printf("     There was species overflow amounting to %s\n", fortran_e93((double)so).c_str());
        }
        if (none[n_acc - 1] > 0) {
            if ((float)none[n_acc - 1] / total > 0.01f) printf("     There were %.1f%% acceptor molecules without any neighbouring donors.\n", (double)(100.0f * (float)none[n_acc - 1] / total));
            else printf("     There were %lld acceptor molecules without neighbouring donors (less than 1%%).\n", none[n_acc - 1]);
        } else printf("     There were no acceptor molecules without neighbouring donors.\n");
        printf("   Detailed statistics written to '%sacceptor%d_speciation_statistics.dat'\n", r.prefix.c_str(), n_acc);
        fprintf(f, "\n The donors of this analysis were:\n");
        for (int n = 0; n < ND; ++n) {
            const SpecSide &D = don[n];
            if (D.atom.size() == 1) {
                if (ND == 1) fprintf(f, " Atom %d (%s) of molecule type %d (%s) was the sole donor.\n", D.atom[0], elem2(D.type, D.atom[0]).c_str(), D.type, formula(D.type).c_str());
                else fprintf(f, "   Donor #%d was atom %d (%s) of molecule type %d (%s).\n", n + 1, D.atom[0], elem2(D.type, D.atom[0]).c_str(), D.type, formula(D.type).c_str());
            } else if (atoms_grouped) {
                fprintf(f, "   Donor #%d was molecule type %d (%s), with the following atom groups:\n", n + 1, D.type, formula(D.type).c_str());
                describe_groups(f, D, "donor");
            } else {
                fprintf(f, "   Donor #%d was molecule type %d (%s), with the following atom indices:\n", n + 1, D.type, formula(D.type).c_str());
                for (size_t i = 0; i < D.atom.size(); ++i) fprintf(f, "     Atom index %d (%s)\n", D.atom[i], elem(D.type, D.atom[i]).c_str());
            }
        }
        fclose(f);
        if (ft) fclose(ft);
    }
    return 0;
}

void write_simple_sumrules(Run &r)                            // Distribution:43-67
{
    const std::string path = r.path_in + "prealpha_simple.inp";
    if (access(path.c_str(), F_OK) == 0) r.notice(114, "overwriting existing 'prealpha_simple.inp'");      // NOTICE 114, e.g. Distribution:48-49
    FILE *f = fopen(path.c_str(), "w");
    if (!f) { r.error(46, "cannot write '" + path + "'"); return; }
    fprintf(f, "pdf %zu\n", r.types.size());
    for (size_t n = 1; n <= r.types.size(); ++n) fprintf(f, "+0 +0 +1 %zu -1\n", n);
    fprintf(f, "maxdist_optimize\nsubtract_uniform F\nweigh_charge T\nsampling_interval 1\nbin_count_a 300\nbin_count_b 1\nquit\n");
    fclose(f);
    if (r.verbose) printf(" File 'prealpha_simple.inp' written.\n");
}

// write_simple_intramolecular_distances / write_simple_intermolecular_distances, Distances:178-284.  The
// intermolecular variant sets both_available from F alone (the reference tests "F" twice, Distances:252-253).
bool write_simple_distances(Run &r, bool inter)
{
    const std::string path = r.path_in + "prealpha_simple.inp";
    if (access(path.c_str(), F_OK) == 0) r.notice(114, "overwriting existing 'prealpha_simple.inp'");      // NOTICE 114, e.g. Distribution:48-49
    FILE *f = fopen(path.c_str(), "w");
    if (!f) { r.error(46, "cannot write '" + path + "'"); return false; }
    bool h_av = false, f_av = false, both = false;
    if (inter) {
        h_av = atoms_of_element(r, "H", 0).size() > 1;
        f_av = atoms_of_element(r, "F", 0).size() > 1;
        // give_number_of_specific_atoms counts atoms per molecule type summed over types, not instances
        both = f_av;
    } else {
        for (size_t t = 0; t < r.types.size(); ++t) {
            const size_t nh = atoms_of_element(r, "H", (int)t + 1).size(), nf = atoms_of_element(r, "F", (int)t + 1).size();
            if (nh > 1) h_av = true;
            if (nf > 1) f_av = true;
            if (nh > 1 && nf > 1) both = true;
        }
    }
    const int n = (h_av ? 1 : 0) + (f_av ? 1 : 0) + (both ? 2 : 0);
    fprintf(f, "%s %d\n", inter ? "inter" : "intra", n);
    if (n == 0) { fclose(f); return false; }
    if (h_av) fprintf(f, "H H\n");
    if (f_av) fprintf(f, "F F\n");
    if (both) fprintf(f, "H F\nF H\n");
    fprintf(f, "exponential_weight T\nffc T\nstdev T\nnsteps %d\ncutoff_optimize\nquit\n", r.nsteps > 10 ? 10 : r.nsteps);
    fclose(f);
    if (r.verbose) printf(" File 'prealpha_simple.inp' written.\n");
    return true;
}

std::string fortran_f0_3(float v)                             // F0.3: no leading zero for |v| < 1
{
    char buf[64];
    snprintf(buf, sizeof buf, "%.3f", (double)v);
    std::string t = buf;
    if (t.rfind("0.", 0) == 0) t.erase(0, 1);
    else if (t.rfind("-0.", 0) == 0) t.erase(1, 1);
    return t;
}

int perform_contact_distance(Run &r, int timestep, int type_in)       // Debug:586-623
{
    const int ntypes = (int)r.types.size();
    if (type_in > ntypes) { r.error(33, "molecule type index does not exist"); return 33; }
    std::vector<pa_type> types; pa_traj traj;
    make_traj(r, types, traj);
    auto print_distances = [&](int type) -> int {
        pa_contact_result res{};
        const int rc = pa_contact_distance(&traj, timestep, type, &r.exec, &res);
        if (rc) { r.gpu_failed = true; r.error(rc, std::string("GPU contact distance failed: ") + pa_last_error()); return rc; }
        const char *ind = type_in < 1 ? "  " : "";
        printf("%s   Largest intramolecular distance:  %s\n", ind, fortran_f0_3(res.largest_intramolecular).c_str());
        printf("%s   Smallest intramolecular distance: %s\n", ind, fortran_f0_3(res.smallest_intramolecular).c_str());
        printf("%s   Smallest intermolecular distance: %s\n", ind, fortran_f0_3(res.smallest_intermolecular).c_str());
        return 0;
    };
    if (type_in < 1) {
        printf(" Calculating intra- and intermolecular contact distances at timestep %d for all molecule types.\n", timestep);
        for (int t = 1; t <= ntypes; ++t) {
            printf("   Molecule type index %d out of %d.\n", t, ntypes);
            const int rc = print_distances(t);
            if (rc) return rc;
        }
        return 0;
    }
    printf(" Calculating intra- and intermolecular contact distances for molecule type %d at timestep %d.\n", type_in, timestep);
    return print_distances(type_in);
}

void write_simple_charge_arm(Run &r, bool normalise)          // Distribution:69-96
{
    const std::string path = r.path_in + "prealpha_simple.inp";
    if (access(path.c_str(), F_OK) == 0) r.notice(114, "overwriting existing 'prealpha_simple.inp'");      // NOTICE 114, e.g. Distribution:48-49
    FILE *f = fopen(path.c_str(), "w");
    if (!f) { r.error(46, "cannot write '" + path + "'"); return; }
    fprintf(f, "charge_arm %zu\n", r.types.size());
    for (size_t n = 1; n <= r.types.size(); ++n) fprintf(f, "+0 +0 +1 %zu\n", n);
    fprintf(f, "maxdist_optimize\nsubtract_uniform F\nsampling_interval 1\nbin_count_a 100\nbin_count_b 100\n");
    if (normalise) fprintf(f, "normalise_clm T\n");
    fprintf(f, "quit\n");
    fclose(f);
    if (r.verbose) printf(" File 'prealpha_simple.inp' written.\n");
}

const char *kOtherModuleKeywords[] = {
    "show_settings", "print_atomic_masses", "print_atomic_charges", "print_dipole_statistics", "show_drude",
    "switch_to_com", "barycentric", "barycenter", "barycentre", "dump_snapshot", "dump_snapshot_simple", "dump_split",
    "dump_split_simple", "dump_single", "dump_cut", "dump_dimers", "dump_neighbour_traj", "dump_example",
    "dump_example_simple", "dump_full_gro", "dump_full_gro_simple", 
    "remove_drudes", "remove_drudes_simple", "remove_cores", "remove_cores_simple", "drude_temp", "drude_temp_simple",
    "temperature", "temperature_simple", "gyradius", "gyradius_simple", "jump_velocity", "jump_velocity_simple",
    "convert", "convert_simple", "convert_coc", "convert_coc_simple", "unwrap_full", "unwrap_full_simple",
    "correlation", "autocorrelation", "dihedral", "reorientation", "rmm-vcf", "conductivity", "velocity",
    "conductivity_simple", "diffusion", "diffusion_simple", "cross_diffusion", "cluster_simple",
    "speciation", "speciation_simple", "charge_arm_simple", "clm_simple", "cubic_box_edge_simple", "time_scaling",
    "time_scaling_factor", "error_output", "time_output", "set_threads", "set_threads_simple", "parallel_operation",
    "parallel_execution", "print_atom_masses", "print_atomic_weights", "print_atom_weights", "print_atom_charges",
    "print_dipole_properties", "show_drude_settings", "show_drudes", "drude_temperature", "remove_drude", "remove_core",
    "dump_slab", "dump_slab_simple", "distance_simple", "distances_simple", "DEBUG",
};

}  // namespace

// test hook: the trajectory readers' decimal -> REAL(4) conversion (must equal strtof bit for bit)
extern "C" float pa_debug_parse_coordinate(const char *s, int32_t *consumed)
{
    const char *e = s;
    const float f = parse_coordinate(s, &e);
    if (consumed) *consumed = (int32_t)(e - s);
    return f;
}

static int run_general_input_impl(const char *general_inp_path, const pa_exec *exec, const int32_t *dbg_conns, int64_t dbg_nconns);

extern "C" int pa_run_general_input(const char *general_inp_path, const pa_exec *exec)
{
    return run_general_input_impl(general_inp_path, exec, nullptr, 0);
}

// Test seam for the host-side bookkeeping of the `speciation` keyword: the same front end, but the connections of every
// (speciation keyword, step, acceptor type) come from the caller's table instead of pa_neighbour_pairs, so that the species
// recognition and the writers can be checked against the reference binary's files on a machine without a GPU (the tests
// fill the table from the oracle's neighbour loops).  Rows of 8 int32: ordinal of the speciation keyword in the input
// (0-based), step, acceptor type (n_acceptor), acceptor molecule, donor type (n_donor), donor molecule, donor atom entry,
// acceptor atom entry (all 1-based).  Nothing in the product calls this; every other keyword behaves as in
// pa_run_general_input and still needs the GPU.
extern "C" int pa_debug_run_general_input_with_connections(const char *general_inp_path, const pa_exec *exec, const int32_t *conns, int64_t nconns)
{
    if (!conns || nconns < 0) return PA_ERR_BAD_ARG;
    return run_general_input_impl(general_inp_path, exec, conns, nconns);
}

static int run_general_input_impl(const char *general_inp_path, const pa_exec *exec, const int32_t *dbg_conns, int64_t dbg_nconns)
{
    if (!general_inp_path) return PA_ERR_BAD_ARG;
    const auto t_start = std::chrono::steady_clock::now();
    const bool dbg_t = getenv("PA_DEBUG_TIMING") != nullptr;
    auto mark = [&](const char *what) {
        if (dbg_t) fprintf(stderr, "[prealpha_b200] cli %-22s %9.1f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
    };
    Run r;
    if (exec) r.exec = *exec;
    if (r.exec.shard_count < 1) r.exec.shard_count = 1;
    r.general_path = general_inp_path;
    r.dbg_conns = dbg_conns; r.dbg_nconns = dbg_nconns;
    // CUDA context creation and the load of the kernel image take ~1 s in a fresh process: start them now, on a helper
    // thread, so that they overlap with reading the input files and the trajectory instead of delaying the first analysis
    // (every device the run may use, one helper thread each: contexts are created concurrently)
    const int warm_device = r.exec.device, warm_n = r.exec.ndevices;
    std::thread warm([warm_device, warm_n] {
        const int usable = pa_device_count();
        if (usable < 1) return;
        const int n = warm_n < 0 ? usable - warm_device : std::max(1, warm_n);
        std::vector<std::thread> th;
        for (int d = warm_device; d < std::min(usable, warm_device + n); ++d)
            th.emplace_back([d] { if (cudaSetDevice(d) == cudaSuccess) { cudaFree(nullptr); pa_kernels_selfcheck(); } });
        for (auto &t : th) t.join();
    });
    struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } warm_joiner{ warm };
    std::vector<std::string> lines;
    if (!read_lines(r.general_path, lines)) { r.error(5, "general input file '" + r.general_path + "' not found"); return r.errors; }
    std::vector<std::vector<std::string>> toks;        // non-blank records, list-directed READ skips blank ones
    for (auto &l : lines) { auto t = tokens_of(l); if (!t.empty()) toks.push_back(t); }
    if (toks.size() < 5) { r.error(5, "general input file has fewer than five header lines"); return r.errors; }
    r.traj_file = toks[0][0]; r.mol_file = toks[1][0]; r.path_traj = toks[2][0]; r.path_in = toks[3][0]; r.path_out = toks[4][0];
    printf(" reading file '%s'\n", r.general_path.c_str());
    std::vector<std::vector<std::string>> body(toks.begin() + 5, toks.end());
    // whole-file switches (first occurrence before 'quit'), Main:131-357
    bool b;
    if (auto *t = find_switch(body, {"sequential_read", "read_sequential"})) if (t->size() > 1 && parse_logical((*t)[1], b)) r.read_sequential = b;
    bool type_given = false;
    if (auto *t = find_switch(body, {"trajectory_type"})) {
        type_given = true;
        if (t->size() > 1 && ((*t)[1].substr(0, 3) == "lmp" || (*t)[1].substr(0, 3) == "xyz")) { r.traj_type = (*t)[1].substr(0, 3); r.box_given = (r.traj_type == "lmp"); }
        else r.error(51, "unknown trajectory format");
    }
    if (!type_given) {
        const std::string ext = r.traj_file.size() >= 4 ? r.traj_file.substr(r.traj_file.size() - 4) : "";
        if (ext == ".lmp" || ext == ".LMP") { r.traj_type = "lmp"; r.box_given = true; }
        else if (ext == ".xyz" || ext == ".XYZ") { r.traj_type = "xyz"; r.box_given = false; }
        else r.error(51, "unknown trajectory format");
    }
    if (auto *t = find_switch(body, {"cubic_box_edge", "cubic_box"})) {
        float lo, hi;
        if (t->size() > 2 && parse_real4((*t)[1], lo) && parse_real4((*t)[2], hi)) {
            if (lo > hi) r.error(93, "lower box bound > upper box bound");
            else { if (r.box_given) r.error(92, "box volume specified twice"); set_cubic_box(r, lo, hi); }
        } else r.error(19, "cannot read cubic_box_edge");
    }
    if (auto *t = find_switch(body, {"wrap_trajectory"})) if (t->size() > 1 && parse_logical((*t)[1], b)) r.wrap = b;
    if (auto *t = find_switch(body, {"unwrap_trajectory"})) if (t->size() > 1 && parse_logical((*t)[1], b)) r.unwrap = b;
    if (r.unwrap && r.wrap) {                                    // Main:309-313, error 151
        r.error(151, "wrap_trajectory and unwrap_trajectory are mutually exclusive - both switched off");
        r.wrap = r.unwrap = false;
    }
    if (r.unwrap && r.read_sequential) { r.error(152, "unwrap_trajectory needs the whole trajectory in RAM"); r.unwrap = false; }
    if (r.read_sequential) {
        // the distribution analysis streams the file batch by batch; the other analyses of this program (distance,
        // contact_distance) and the wrap / unwrap pre-passes work on the trajectory in RAM
        r.stream = !r.wrap && !r.unwrap;
        for (const auto &t : body) {
            if (t.empty()) continue;
            if (t[0] == "quit" || t[0] == "exit") break;
            if (t[0] == "distance" || t[0] == "distances" || t[0] == "distance_simple" || t[0] == "distances_simple" ||
                t[0] == "contact_distance" || t[0] == "contact_distance_simple" || t[0] == "cluster" || t[0] == "speciation") r.stream = false;
        }
        printf(r.stream ? " sequential_read T: the distribution analysis streams the trajectory from the file, batch by batch.\n"
                        : " sequential_read T: this input needs the whole trajectory in RAM; it is loaded once.\n");
    }
    if (!read_molecular_input(r)) return r.errors;
    mark("inputs read");
    if (!load_trajectory(r)) return r.errors;
    mark("trajectory loaded");
    if ((r.wrap || r.unwrap) && !r.box_given) { r.error(41, "box volume required for wrapping"); return r.errors; }
    if (r.wrap) {                                                // Molecular:4801-4804
        printf(" Wrapping centres of mass of each specified molecule into box *now*.\n");
        wrap_full(r);
    }
    if (r.unwrap) {                                              // Molecular:4806-4816 (developers version unwraps twice)
        printf(" Preparing a clean trajectory before unwrapping...\n");
        wrap_full(r);
        printf(" UNwrapping centres of mass of each specified molecule *now*.\n");
        unwrap_full(r);
        unwrap_full(r);
    }
    // body, Main:2882-3666
    // Main:2869: only the header of a lammps trajectory says whether it holds positions or velocities (Molecular:4861-4870)
    if (r.traj_type == "xyz") r.error(55, "unknown trajectory format (velocities or Cartesian coordinates?) - please check results and input");
    int nline = 5;
    for (const auto &t : body) {
        ++nline;
        if (nline - 5 > 500) break;                          // MAXITERATIONS
        std::string k = t[0];
        if (k == "exit") k = "quit";
        if (r.verbose) printf(" Line %d: '%s'\n", nline, k.c_str());
        if (k == "quit") { printf(" exiting analysis.\n"); break; }
        if (k == "distribution") {
            if (!r.box_given) { r.error(41, "box volume required for this analysis"); continue; }
            if (t.size() < 2) { r.error(19, "cannot read general input line"); break; }
            printf(" Distribution module invoked.\n");
            mark("distribution begins");
            perform_distribution(r, t[1]);
            mark("distribution done");
        } else if (k == "distribution_simple") {
            if (!r.box_given) { r.error(41, "box volume required for this analysis"); continue; }
            write_simple_sumrules(r);
            printf(" Distribution module invoked.\n");
            perform_distribution(r, "prealpha_simple.inp");
        } else if (k == "charge_arm_simple" || k == "clm_simple" || k == "CLM_simple" || k == "charge_lever_moment_simple") {   // Main:3518-3535
            const bool clm = k != "charge_arm_simple";
            printf(" Distribution module invoked - simple mode:\n");
            printf(clm ? " CLM distribution. Requires charges to be initialised.\n (charge arm with charge lever moment correction)\n"
                       : " Charge Arm distribution. Requires charges to be initialised.\n");
            write_simple_charge_arm(r, clm);
            perform_distribution(r, "prealpha_simple.inp");
        } else if (k == "contact_distance" || k == "contact_distance_simple") {   // Main:3066-3086
            if (!r.box_given) { r.error(41, "box volume required for this analysis"); continue; }
            int step = 1, type = -1;
            if (k == "contact_distance" && !(t.size() > 2 && parse_int(t[1], step) && parse_int(t[2], type))) { r.error(19, "cannot read general input line"); break; }
            if (step < 1) { r.error(57, "timestep out of range"); step = 1; }                 // check_timestep, Debug:46-56
            else if (step > r.nsteps) { r.error(57, "timestep out of range"); step = r.nsteps; }
            perform_contact_distance(r, step, type);
        } else if (k == "distance_simple" || k == "distances_simple") {      // Main:3394-3409
            printf(" Calculating, where possible, intra- and intermolecular distances for H and F.\n");
            for (int inter = 0; inter < 2; ++inter) {
                if (write_simple_distances(r, inter != 0)) {
                    printf(" %smolecular distance calculation, simple mode:\n", inter ? "Inter" : "Intra");
                    perform_distance(r, "prealpha_simple.inp");
                } else printf(" Can't do %smolecular distances. Do manually if necessary.\n", inter ? "inter" : "intra");
            }
        } else if (k == "distance" || k == "distances") {                    // Main:3379-3390
            if (t.size() < 2) { r.error(19, "cannot read general input line"); break; }
            printf(" distance module invoked.\n");
            perform_distance(r, t[1]);
        } else if (k == "cluster") {                                          // Main: cluster keyword -> perform_cluster_analysis (Cluster:1931)
            if (!r.box_given) { r.error(41, "box volume required for this analysis"); continue; }
            if (t.size() < 2) { r.error(19, "cannot read general input line"); break; }
            printf(" Cluster module invoked.\n");
            perform_cluster(r, t[1]);
        } else if (k == "speciation") {                                       // Main: speciation keyword -> perform_speciation_analysis (Speciation:3230)
            if (!r.box_given) { r.error(41, "box volume required for this analysis"); continue; }
            if (t.size() < 2) { r.error(19, "cannot read general input line"); break; }
            printf(" Speciation module invoked.\n");
            perform_speciation(r, t[1]);
        } else if (k == "set_prefix") {
            if (t.size() < 2) { r.error(19, "cannot read general input line"); break; }
            r.prefix = t[1];
            while (!r.prefix.empty() && r.prefix.front() == ' ') r.prefix.erase(r.prefix.begin());
            printf(" prefix set to '%s'\n", r.prefix.c_str());
        } else if (k == "verbose_output") {
            if (t.size() > 1 && parse_logical(t[1], b)) r.verbose = b; else { r.error(19, "cannot read general input line"); break; }
        } else if (k == "sequential_read" || k == "read_sequential" || k == "wrap_trajectory" || k == "unwrap_trajectory" ||
                   k == "trajectory_type" || k == "cubic_box_edge" || k == "cubic_box" || k == "extra_velocity" || k == "extra_columns") {
            if (r.verbose) printf(" skip line (%s)\n", k.c_str());
        } else if (k[0] == '#' || k[0] == '!') {
            if (r.verbose) printf(" (is commented out)\n");
        } else {
            bool known = false;
            for (const char *o : kOtherModuleKeywords) if (k == o) { known = true; break; }
            if (known) printf(" keyword '%s' belongs to a prealpha module outside the accelerated path - skipped.\n", k.c_str());
            else r.error(20, "unknown keyword '" + k + "' in line " + std::to_string(nline));
        }
    }
    printf("\n Encountered %d errors/warnings globally.\n", r.errors);
    mark("all keywords done");
    return r.gpu_failed ? PA_ERR_CUDA : r.errors;
}

