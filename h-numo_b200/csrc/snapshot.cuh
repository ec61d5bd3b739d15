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

}  // extern "C"
