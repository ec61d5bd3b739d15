// snapshot.cuh -- host-only reader / writer of the reference's text snapshots and the restart conversion.
//
//   writer: the `mlswe####` files of src/diagnostics.F90:73-91 -- nlayers (i4), npoin (i10), dt, dt_btp, then one value per line
//           in Fortran d23.16: coord(1:2,:), pb, pb*ub, pb*vb, h(:,k), u(:,k), v(:,k), interface elevation(:,k) for all layers, zbot
//           (h = alpha_k/g dp_k, u = (u dp)/dp, elevation_k = zbot + sum_{j>=k} h_j: src/diagnostics.F90:24-45)
//   reader: load_data_mlswe as used by read_mlswe (src/mod_restart.F90:88-160)
//   restart conversion: restart_mlswe (src/mod_restart.F90:15-66): qb = (pb, pb - pbprime, pb ub, pb vb), dp = g/alpha h,
//           (u dp, v dp), dp' = dp / (sum_k dp_k / pbprime), u' = u - ub, v' = v - vb
// so that a GPU run can start from a reference snapshot (hnumo_upload_state) and the reference can restart from a GPU run.
// No device code: these entry points work without a GPU.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace hn {

// Fortran d23.16: "  0.1234567890123456D+03" (three-digit exponents drop the letter: "0.1234567890123456+100")
inline void fmt_d23_16(double x, char* out /* >= 32 bytes */) {
    if (x == 0.0 || !std::isfinite(x)) {
        if (x == 0.0) snprintf(out, 32, "%23s", std::signbit(x) ? "-0.0000000000000000D+00" : "0.0000000000000000D+00");
        else snprintf(out, 32, "%23s", std::isnan(x) ? "NaN" : (x > 0 ? "Infinity" : "-Infinity"));
        return;
    }
    char buf[40];
    snprintf(buf, sizeof(buf), "%.15e", fabs(x));   // d.ddddddddddddddde+XX : 16 significant digits, correctly rounded
    char digits[17]; int nd = 0;
    int ex = 0;
    for (const char* p = buf; *p; ++p) {
        if (*p == 'e') { ex = atoi(p + 1); break; }
        if (*p >= '0' && *p <= '9') digits[nd++] = *p;
    }
    digits[nd] = 0;
    ex += 1;   // 0.dddd form
    char body[48];
    if (abs(ex) < 100) snprintf(body, sizeof(body), "%s0.%sD%c%02d", x < 0 ? "-" : "", digits, ex < 0 ? '-' : '+', abs(ex));
    else snprintf(body, sizeof(body), "%s0.%s%c%03d", x < 0 ? "-" : "", digits, ex < 0 ? '-' : '+', abs(ex));
    snprintf(out, 32, "%23s", body);
}

inline bool parse_fortran_real(const char* s, double* v) {
    char buf[64]; int n = 0;
    while (*s == ' ' || *s == '\t') ++s;
    for (; *s && *s != '\n' && *s != '\r' && n < 62; ++s) {
        char c = *s;
        if (c == 'D' || c == 'd') c = 'E';
        // exponent without a letter: a sign that follows a digit
        if ((c == '+' || c == '-') && n > 0 && ((buf[n - 1] >= '0' && buf[n - 1] <= '9') || buf[n - 1] == '.')) buf[n++] = 'E';
        buf[n++] = c;
    }
    buf[n] = 0;
    if (n == 0) return false;
    char* end = nullptr;
    *v = strtod(buf, &end);
    return end != buf;
}

}  // namespace hn

extern "C" {

// Write a reference-format text snapshot (src/diagnostics.F90:73-91).  coord: (2,npoin) or null (zeros are written);
// q_df(3,npoin,nlayers), qb_df(4,npoin) in the reference layouts.  Returns 0, or -6 on an I/O error.
int hnumo_snapshot_write(const char* path, int32_t nlayers, int64_t npoin, double dt, double dt_btp, const double* coord,
                         const double* q_df, const double* qb_df, const double* zbot_df, const double* alpha_mlswe, double gravity) {
    if (!path || !q_df || !qb_df || !zbot_df || !alpha_mlswe) return -2;
    FILE* f = fopen(path, "w");
    if (!f) { hn::set_error("hnumo_snapshot_write", "cannot open the file"); return -6; }
    char b[40];
    auto put = [&](double x) { hn::fmt_d23_16(x, b); fputs(b, f); fputc('\n', f); };
    fprintf(f, "%4d\n%10lld\n", (int)nlayers, (long long)npoin);
    put(dt); put(dt_btp);
    for (int64_t i = 0; i < npoin; ++i) { put(coord ? coord[2 * i] : 0.0); put(coord ? coord[2 * i + 1] : 0.0); }
    for (int v : {0, 2, 3})
        for (int64_t i = 0; i < npoin; ++i) put(qb_df[4 * i + v]);
    auto Q = [&](int v, int64_t i, int k) { return q_df[((size_t)k * npoin + i) * 3 + v]; };
    for (int k = 0; k < nlayers; ++k)
        for (int64_t i = 0; i < npoin; ++i) put((alpha_mlswe[k] / gravity) * Q(0, i, k));
    for (int v = 1; v <= 2; ++v)
        for (int k = 0; k < nlayers; ++k)
            for (int64_t i = 0; i < npoin; ++i) put(Q(v, i, k) / Q(0, i, k));
    // interface elevations: elevation_k = zbot + sum_{j >= k} h_j (src/diagnostics.F90:31-45)
    std::vector<double> elev((size_t)nlayers * npoin);
    for (int64_t i = 0; i < npoin; ++i) {
        double z = zbot_df[i];
        for (int k = nlayers - 1; k >= 0; --k) { z += (alpha_mlswe[k] / gravity) * Q(0, i, k); elev[(size_t)k * npoin + i] = z; }
    }
    for (size_t j = 0; j < elev.size(); ++j) put(elev[j]);
    for (int64_t i = 0; i < npoin; ++i) put(zbot_df[i]);
    int rc = ferror(f) ? -6 : 0;
    fclose(f);
    if (rc) hn::set_error("hnumo_snapshot_write", "write error");
    return rc;
}

// Read the header of a snapshot: nlayers, npoin, dt, dt_btp.  Returns 0 or <0.
int hnumo_snapshot_info(const char* path, int32_t* nlayers, int64_t* npoin, double* dt, double* dt_btp) {
    FILE* f = fopen(path, "r");
    if (!f) { hn::set_error("hnumo_snapshot_info", "cannot open the file"); return -6; }
    char line[128];
    long long np = 0; int nl = 0; double a = 0, b = 0;
    bool ok = fgets(line, sizeof(line), f) && sscanf(line, "%d", &nl) == 1 && fgets(line, sizeof(line), f) && sscanf(line, "%lld", &np) == 1 &&
              fgets(line, sizeof(line), f) && hn::parse_fortran_real(line, &a) && fgets(line, sizeof(line), f) && hn::parse_fortran_real(line, &b);
    fclose(f);
    if (!ok) { hn::set_error("hnumo_snapshot_info", "malformed header"); return -7; }
    if (nlayers) *nlayers = nl; if (npoin) *npoin = np; if (dt) *dt = a; if (dt_btp) *dt_btp = b;
    return 0;
}

// Read a snapshot and convert it to the prognostic arrays exactly as restart_mlswe does (src/mod_restart.F90:15-66).
// pbprime_df(npoin), alpha_mlswe(nlayers) are the statics of the run that restarts.  Outputs in the reference layouts;
// coord_out (2,npoin) may be null.  Returns 0 or <0 (-7 malformed file, -8 size mismatch).
int hnumo_snapshot_read_restart(const char* path, int32_t nlayers, int64_t npoin, const double* pbprime_df, const double* alpha_mlswe,
                                double gravity, double* q_df, double* qb_df, double* qprime_df, double* coord_out) {
    if (!path || !pbprime_df || !alpha_mlswe || !q_df || !qb_df || !qprime_df) return -2;
    FILE* f = fopen(path, "r");
    if (!f) { hn::set_error("hnumo_snapshot_read_restart", "cannot open the file"); return -6; }
    char line[128];
    long long np = 0; int nl = 0;
    if (!(fgets(line, sizeof(line), f) && sscanf(line, "%d", &nl) == 1 && fgets(line, sizeof(line), f) && sscanf(line, "%lld", &np) == 1)) {
        fclose(f); hn::set_error("hnumo_snapshot_read_restart", "malformed header"); return -7;
    }
    if (nl != nlayers || np != npoin) { fclose(f); hn::set_error("hnumo_snapshot_read_restart", "nlayers / npoin differ from the run"); return -8; }
    bool ok = true;
    auto get = [&]() { double v = 0.0; if (!(fgets(line, sizeof(line), f) && hn::parse_fortran_real(line, &v))) ok = false; return v; };
    get(); get();   // dt, dt_btp
    for (int64_t i = 0; i < npoin; ++i) { double x = get(), y = get(); if (coord_out) { coord_out[2 * i] = x; coord_out[2 * i + 1] = y; } }
    for (int v : {0, 2, 3})
        for (int64_t i = 0; i < npoin; ++i) qb_df[4 * i + v] = get();
    for (int64_t i = 0; i < npoin; ++i) qb_df[4 * i + 1] = qb_df[4 * i] - pbprime_df[i];
    auto Q = [&](int v, int64_t i, int k) -> double& { return q_df[((size_t)k * npoin + i) * 3 + v]; };
    for (int k = 0; k < nlayers; ++k)
        for (int64_t i = 0; i < npoin; ++i) Q(0, i, k) = (gravity / alpha_mlswe[k]) * get();
    for (int v = 1; v <= 2; ++v)
        for (int k = 0; k < nlayers; ++k)
            for (int64_t i = 0; i < npoin; ++i) Q(v, i, k) = get() * Q(0, i, k);
    fclose(f);   // elevations and zbot are diagnostics: not needed for the restart
    if (!ok) { hn::set_error("hnumo_snapshot_read_restart", "file is shorter than its header says"); return -7; }
    for (int64_t i = 0; i < npoin; ++i) {
        double s = 0.0;
        for (int k = 0; k < nlayers; ++k) s += Q(0, i, k);
        const double ope = s / pbprime_df[i];
        for (int k = 0; k < nlayers; ++k) {
            double* qp = qprime_df + ((size_t)k * npoin + i) * 3;
            qp[0] = Q(0, i, k) / ope;
            qp[1] = Q(1, i, k) / Q(0, i, k) - qb_df[4 * i + 2] / qb_df[4 * i];
            qp[2] = Q(2, i, k) / Q(0, i, k) - qb_df[4 * i + 3] / qb_df[4 * i];
        }
    }
    return 0;
}


// ---- NetCDF (classic, 64-bit offset: CDF-2) snapshots, written and read without a NetCDF library ---------------------------
// hnumo_snapshot_write_nc writes the file diagnostics_nc writes (src/diagnostics_nc.F90:98-165: nf90_create(...,
// nf90_clobber + nf90_64bit_offset)): dimensions time (unlimited), npoin, nlayers, zi = nlayers + 1; variables dt(time),
// dt_btp(time), x, y, pb, pbub, pbvb (npoin), h, u, v (nlayers, npoin), eta (zi, npoin) -- the Fortran dimension order
// (npoin, nlayers) reversed, as every NetCDF file stores it -- with the reference's names and attributes.
// eta(:,1) = pb / pbprime - 1, eta(:,2:nlayers) = interface elevations, eta(:,nlayers+1) = zbot (src/diagnostics_nc.F90:62-85).
}  // extern "C"
namespace hn {
struct NcOut {
    std::vector<unsigned char> b;
    void u32(uint32_t v) { for (int i = 3; i >= 0; --i) b.push_back((unsigned char)(v >> (8 * i))); }
    void u64(uint64_t v) { for (int i = 7; i >= 0; --i) b.push_back((unsigned char)(v >> (8 * i))); }
    void name(const char* s) { size_t n = strlen(s); u32((uint32_t)n); for (size_t i = 0; i < n; ++i) b.push_back((unsigned char)s[i]); while (b.size() % 4) b.push_back(0); }
    void att(const char* nm, const char* val) { name(nm); u32(2 /*NC_CHAR*/); size_t n = strlen(val); u32((uint32_t)n); for (size_t i = 0; i < n; ++i) b.push_back((unsigned char)val[i]); while (b.size() % 4) b.push_back(0); }
};
struct NcVar { const char* name; int ndims; int dims[2]; const char* a1n; const char* a1v; const char* a2n; const char* a2v; uint64_t vsize, begin; };
static inline void nc_put_doubles(FILE* f, const double* v, size_t n) {
    std::vector<unsigned char> buf(8 * std::min<size_t>(n, 1 << 16));
    for (size_t o = 0; o < n; o += (1 << 16)) {
        size_t m = std::min<size_t>(1 << 16, n - o);
        for (size_t i = 0; i < m; ++i) { uint64_t u; memcpy(&u, &v[o + i], 8); for (int k = 0; k < 8; ++k) buf[8 * i + k] = (unsigned char)(u >> (8 * (7 - k))); }
        fwrite(buf.data(), 1, 8 * m, f);
    }
}
}  // namespace hn
extern "C" {

int hnumo_snapshot_write_nc(const char* path, int32_t nlayers, int64_t npoin, double dt, double dt_btp, const double* coord,
                            const double* q_df, const double* qb_df, const double* zbot_df, const double* alpha_mlswe,
                            const double* pbprime_df, double gravity) {
    using namespace hn;
    if (!path || !q_df || !qb_df || !zbot_df || !alpha_mlswe || !pbprime_df) return -2;
    enum { D_TIME = 0, D_NPOIN = 1, D_NL = 2, D_ZI = 3 };
    const uint64_t NP = (uint64_t)npoin, NL = (uint64_t)nlayers;
    NcVar vars[11] = {
        {"dt", 1, {D_TIME, 0}, "name", "Baroclinic time step", "units", "seconds", 8, 0},
        {"dt_btp", 1, {D_TIME, 0}, "name", "Barotropic time step", "units", "seconds", 8, 0},
        {"x", 1, {D_NPOIN, 0}, "name", "cartesian coordinates", "axis", "X", 8 * NP, 0},
        {"y", 1, {D_NPOIN, 0}, "name", "cartesian coordinates", "axis", "Y", 8 * NP, 0},
        {"pb", 1, {D_NPOIN, 0}, "name", "Barotropic pressure pb", "units", "N/m\xc2\xb2", 8 * NP, 0},
        {"pbub", 1, {D_NPOIN, 0}, "name", "Barotropic u-momentum", "units", "kg\xc2\xb7m/s", 8 * NP, 0},
        {"pbvb", 1, {D_NPOIN, 0}, "name", "Barotropic v-momentum", "units", "kg\xc2\xb7m/s", 8 * NP, 0},
        {"h", 2, {D_NL, D_NPOIN}, "name", "Layer thickness", "units", "m", 8 * NP * NL, 0},
        {"u", 2, {D_NL, D_NPOIN}, "name", "Baroclinic u-velocity", "units", "m/s", 8 * NP * NL, 0},
        {"v", 2, {D_NL, D_NPOIN}, "name", "Baroclinic v-velocity", "units", "m/s", 8 * NP * NL, 0},
        {"eta", 2, {D_ZI, D_NPOIN}, "name", "Interface Height Relative to Mean Sea Level", "units", "m", 8 * NP * (NL + 1), 0},
    };
    const char* base = strrchr(path, '/');
    base = base ? base + 1 : path;
    // two passes: the first sizes the header, the second writes it with the data offsets
    NcOut H;
    for (int pass = 0; pass < 2; ++pass) {
        const uint64_t hdr = H.b.size();
        uint64_t off = hdr;
        for (int i = 2; i < 11; ++i) { vars[i].begin = off; off += vars[i].vsize; }     // fixed-size variables, definition order
        vars[0].begin = off; vars[1].begin = off + 8;                                    // then the records: (dt, dt_btp) per record
        H.b.clear();
        H.b.push_back('C'); H.b.push_back('D'); H.b.push_back('F'); H.b.push_back(2);
        H.u32(1);                                                                         // numrecs
        H.u32(0x0A); H.u32(4);
        H.name("time"); H.u32(0); H.name("npoin"); H.u32((uint32_t)NP); H.name("nlayers"); H.u32((uint32_t)NL); H.name("zi"); H.u32((uint32_t)NL + 1);
        H.u32(0x0C); H.u32(3);
        H.att("filename", base); H.att("npoin", "Number of points in the mesh"); H.att("zi", "Number of interfaces");
        H.u32(0x0B); H.u32(11);
        for (int i = 0; i < 11; ++i) {
            const NcVar& v = vars[i];
            H.name(v.name); H.u32((uint32_t)v.ndims);
            for (int d = 0; d < v.ndims; ++d) H.u32((uint32_t)v.dims[d]);
            H.u32(0x0C); H.u32(2); H.att(v.a1n, v.a1v); H.att(v.a2n, v.a2v);
            H.u32(6 /*NC_DOUBLE*/);
            H.u32(v.vsize > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)v.vsize);
            H.u64(v.begin);
        }
    }
    FILE* f = fopen(path, "wb");
    if (!f) { set_error("hnumo_snapshot_write_nc", "cannot open the file"); return -6; }
    fwrite(H.b.data(), 1, H.b.size(), f);
    std::vector<double> t((size_t)NP * (NL + 1));
    auto Q = [&](int v, int64_t i, int k) { return q_df[((size_t)k * npoin + i) * 3 + v]; };
    for (int c = 0; c < 2; ++c) { for (uint64_t i = 0; i < NP; ++i) t[i] = coord ? coord[2 * i + c] : 0.0; nc_put_doubles(f, t.data(), NP); }
    for (int v : {0, 2, 3}) { for (uint64_t i = 0; i < NP; ++i) t[i] = qb_df[4 * i + v]; nc_put_doubles(f, t.data(), NP); }
    for (int k = 0; k < nlayers; ++k) for (uint64_t i = 0; i < NP; ++i) t[(size_t)k * NP + i] = (alpha_mlswe[k] / gravity) * Q(0, i, k);
    nc_put_doubles(f, t.data(), NP * NL);
    for (int v = 1; v <= 2; ++v) {
        for (int k = 0; k < nlayers; ++k) for (uint64_t i = 0; i < NP; ++i) t[(size_t)k * NP + i] = Q(v, i, k) / Q(0, i, k);
        nc_put_doubles(f, t.data(), NP * NL);
    }
    for (uint64_t i = 0; i < NP; ++i) {
        double z = zbot_df[i];
        t[(size_t)NL * NP + i] = z;
        for (int k = nlayers - 1; k >= 1; --k) { z += (alpha_mlswe[k] / gravity) * Q(0, i, k); t[(size_t)k * NP + i] = z; }
        t[i] = qb_df[4 * i] * (pbprime_df[i] > 0.0 ? 1.0 / pbprime_df[i] : 0.0) - 1.0;
    }
    nc_put_doubles(f, t.data(), NP * (NL + 1));
    double rec[2] = {dt, dt_btp};
    nc_put_doubles(f, rec, 2);
    int rc = ferror(f) ? -6 : 0;
    fclose(f);
    if (rc) set_error("hnumo_snapshot_write_nc", "write error");
    return rc;
}

// Read such a file back and rebuild q_df, qb_df, qprime_df like restart_mlswe does from the text snapshot
// (src/mod_restart.F90:15-66); coord_out (2,npoin) may be null.  Returns 0, -6 I/O, -7 not a CDF-1/2 file of this layout,
// -8 nlayers/npoin mismatch.
int hnumo_snapshot_read_nc_restart(const char* path, int32_t nlayers, int64_t npoin, const double* pbprime_df, const double* alpha_mlswe,
                                   double gravity, double* q_df, double* qb_df, double* qprime_df, double* coord_out) {
    using namespace hn;
    if (!path || !pbprime_df || !alpha_mlswe || !q_df || !qb_df || !qprime_df) return -2;
    FILE* f = fopen(path, "rb");
    if (!f) { set_error("hnumo_snapshot_read_nc_restart", "cannot open the file"); return -6; }
    std::vector<unsigned char> hb(1 << 16);
    size_t got = fread(hb.data(), 1, hb.size(), f);
    size_t pos = 0;
    bool ok = got > 8 && hb[0] == 'C' && hb[1] == 'D' && hb[2] == 'F' && (hb[3] == 1 || hb[3] == 2);
    const int ver = ok ? hb[3] : 0;
    auto r32 = [&]() -> uint32_t { if (pos + 4 > got) { ok = false; return 0; } uint32_t v = ((uint32_t)hb[pos] << 24) | ((uint32_t)hb[pos + 1] << 16) | ((uint32_t)hb[pos + 2] << 8) | hb[pos + 3]; pos += 4; return v; };
    auto rname = [&]() -> std::string { uint32_t n = r32(); if (pos + n > got) { ok = false; return ""; } std::string s((const char*)&hb[pos], n); pos += (n + 3) & ~3u; return s; };
    auto skip_atts = [&]() { uint32_t tag = r32(), n = r32(); if (tag == 0) return; for (uint32_t i = 0; i < n && ok; ++i) { rname(); uint32_t ty = r32(), ne = r32(); size_t sz = (ty == 1 || ty == 2) ? 1 : ty == 3 ? 2 : (ty == 4 || ty == 5) ? 4 : 8; pos += ((size_t)ne * sz + 3) & ~(size_t)3; } };
    pos = 4; r32();
    std::vector<std::pair<std::string, uint32_t>> dims;
    { uint32_t tag = r32(), n = r32(); if (tag == 0x0A) for (uint32_t i = 0; i < n && ok; ++i) { std::string nm = rname(); dims.push_back({nm, r32()}); } }
    skip_atts();
    std::map<std::string, uint64_t> begin;
    { uint32_t tag = r32(), n = r32(); if (tag == 0x0B) for (uint32_t i = 0; i < n && ok; ++i) {
          std::string nm = rname(); uint32_t nd = r32(); for (uint32_t d = 0; d < nd; ++d) r32();
          skip_atts(); r32(); r32();
          uint64_t b = r32(); if (ver == 2) b = (b << 32) | r32();
          begin[nm] = b; } }
    auto dimlen = [&](const char* nm) -> int64_t { for (auto& d : dims) if (d.first == nm) return d.second; return -1; };
    if (!ok || !begin.count("pb") || !begin.count("h") || !begin.count("u") || !begin.count("v")) { fclose(f); set_error("hnumo_snapshot_read_nc_restart", "not a NetCDF classic file with the layout of diagnostics_nc"); return -7; }
    if (dimlen("npoin") != npoin || dimlen("nlayers") != nlayers) { fclose(f); set_error("hnumo_snapshot_read_nc_restart", "nlayers / npoin differ from the run"); return -8; }
    std::vector<double> t((size_t)npoin * nlayers);
    auto rd = [&](const char* nm, size_t n) { std::vector<unsigned char> raw(8 * n); if (fseek(f, (long)begin[nm], SEEK_SET) != 0 || fread(raw.data(), 1, raw.size(), f) != raw.size()) { ok = false; return; }
                                             for (size_t i = 0; i < n; ++i) { uint64_t u = 0; for (int k = 0; k < 8; ++k) u = (u << 8) | raw[8 * i + k]; memcpy(&t[i], &u, 8); } };
    const size_t NP = (size_t)npoin;
    if (coord_out && begin.count("x") && begin.count("y")) { rd("x", NP); for (size_t i = 0; i < NP; ++i) coord_out[2 * i] = t[i]; rd("y", NP); for (size_t i = 0; i < NP; ++i) coord_out[2 * i + 1] = t[i]; }
    const char* bn[3] = {"pb", "pbub", "pbvb"}; const int bv[3] = {0, 2, 3};
    for (int j = 0; j < 3; ++j) { rd(bn[j], NP); for (size_t i = 0; i < NP; ++i) qb_df[4 * i + bv[j]] = t[i]; }
    for (size_t i = 0; i < NP; ++i) qb_df[4 * i + 1] = qb_df[4 * i] - pbprime_df[i];
    auto Q = [&](int v, size_t i, int k) -> double& { return q_df[((size_t)k * npoin + i) * 3 + v]; };
    rd("h", NP * nlayers);
    for (int k = 0; k < nlayers; ++k) for (size_t i = 0; i < NP; ++i) Q(0, i, k) = (gravity / alpha_mlswe[k]) * t[(size_t)k * NP + i];
    rd("u", NP * nlayers);
    for (int k = 0; k < nlayers; ++k) for (size_t i = 0; i < NP; ++i) Q(1, i, k) = t[(size_t)k * NP + i] * Q(0, i, k);
    rd("v", NP * nlayers);
    for (int k = 0; k < nlayers; ++k) for (size_t i = 0; i < NP; ++i) Q(2, i, k) = t[(size_t)k * NP + i] * Q(0, i, k);
    fclose(f);
    if (!ok) { set_error("hnumo_snapshot_read_nc_restart", "file is shorter than its header says"); return -7; }
    for (size_t i = 0; i < NP; ++i) {
        double s = 0.0;
        for (int k = 0; k < nlayers; ++k) s += Q(0, i, k);
        const double ope = s / pbprime_df[i];
        for (int k = 0; k < nlayers; ++k) {
            double* qp = qprime_df + ((size_t)k * npoin + i) * 3;
            qp[0] = Q(0, i, k) / ope;
            qp[1] = Q(1, i, k) / Q(0, i, k) - qb_df[4 * i + 2] / qb_df[4 * i];
            qp[2] = Q(2, i, k) / Q(0, i, k) - qb_df[4 * i + 3] / qb_df[4 * i];
        }
    }
    return 0;
}

}  // extern "C"
