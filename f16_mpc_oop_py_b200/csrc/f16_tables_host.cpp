// f16_tables_host.cpp -- host side of the aero database: read the packed blob or the reference's C/*.dat text
// files into the canonical payload, checksum it, and lay out the device images described in f16_tables.h.
//
// Reference data formats: whitespace-separated decimal text, column-major with alpha fastest, one file per
// table (C/hifi_F16_AeroData.c:136-145 and the 42 sibling loaders; breakpoints :7-105).
#include "f16_tables_host.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <string>
#include <vector>

#include "f16_tables.h"

namespace f16 {

// ---- sha256 (FIPS 180-4), small and self-contained -----------------------------------------------------
namespace {
struct Sha256 {
  uint32_t h[8];
  uint8_t buf[64];
  uint64_t len;
  size_t fill;
};
const uint32_t KK[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
void sha_block(Sha256& s, const uint8_t* p) {
  uint32_t w[64];
  for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
  for (int i = 16; i < 64; i++) {
    uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
    uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t a = s.h[0], b = s.h[1], c = s.h[2], d = s.h[3], e = s.h[4], f = s.h[5], g = s.h[6], hh = s.h[7];
  for (int i = 0; i < 64; i++) {
    uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25), ch = (e & f) ^ (~e & g);
    uint32_t t1 = hh + S1 + ch + KK[i] + w[i];
    uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22), mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  s.h[0] += a; s.h[1] += b; s.h[2] += c; s.h[3] += d; s.h[4] += e; s.h[5] += f; s.h[6] += g; s.h[7] += hh;
}
void sha_digest(const void* data, size_t n, uint8_t out[32]) {
  Sha256 s = {{0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19}, {0}, 0, 0};
  const uint8_t* p = (const uint8_t*)data;
  size_t full = n / 64;
  for (size_t i = 0; i < full; i++) sha_block(s, p + 64 * i);
  uint8_t tail[128] = {0};
  size_t rem = n - full * 64;
  memcpy(tail, p + full * 64, rem);
  tail[rem] = 0x80;
  size_t tl = rem + 1 + 8 <= 64 ? 64 : 128;
  uint64_t bits = (uint64_t)n * 8;
  for (int i = 0; i < 8; i++) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
  for (size_t i = 0; i < tl; i += 64) sha_block(s, tail + i);
  for (int i = 0; i < 8; i++) {
    out[4 * i] = s.h[i] >> 24; out[4 * i + 1] = s.h[i] >> 16; out[4 * i + 2] = s.h[i] >> 8; out[4 * i + 3] = s.h[i];
  }
}

struct TableFile {
  const char* file;
  int n;
};
// order = enum F16CanonTable
const TableFile BREAKS[5] = {{"ALPHA1.dat", 20}, {"ALPHA2.dat", 14}, {"BETA1.dat", 19}, {"DH1.dat", 5}, {"DH2.dat", 3}};
const TableFile TABLES[FT_COUNT] = {
    {"CX0120_ALPHA1_BETA1_DH1_201.dat", 1900}, {"CZ0120_ALPHA1_BETA1_DH1_301.dat", 1900},
    {"CM0120_ALPHA1_BETA1_DH1_101.dat", 1900}, {"CN0120_ALPHA1_BETA1_DH2_501.dat", 1140},
    {"CL0120_ALPHA1_BETA1_DH2_601.dat", 1140}, {"CY0320_ALPHA1_BETA1_401.dat", 380},
    {"CY0720_ALPHA1_BETA1_405.dat", 380},      {"CN0720_ALPHA1_BETA1_503.dat", 380},
    {"CL0720_ALPHA1_BETA1_603.dat", 380},      {"CY0620_ALPHA1_BETA1_403.dat", 380},
    {"CN0620_ALPHA1_BETA1_504.dat", 380},      {"CL0620_ALPHA1_BETA1_604.dat", 380},
    {"CX0820_ALPHA2_BETA1_202.dat", 266},      {"CZ0820_ALPHA2_BETA1_302.dat", 266},
    {"CM0820_ALPHA2_BETA1_102.dat", 266},      {"CY0820_ALPHA2_BETA1_402.dat", 266},
    {"CN0820_ALPHA2_BETA1_502.dat", 266},      {"CL0820_ALPHA2_BETA1_602.dat", 266},
    {"CY0920_ALPHA2_BETA1_404.dat", 266},      {"CN0920_ALPHA2_BETA1_505.dat", 266},
    {"CL0920_ALPHA2_BETA1_605.dat", 266},      {"CX1120_ALPHA1_204.dat", 20},
    {"CZ1120_ALPHA1_304.dat", 20},             {"CM1120_ALPHA1_104.dat", 20},
    {"CY1220_ALPHA1_408.dat", 20},             {"CY1320_ALPHA1_406.dat", 20},
    {"CN1320_ALPHA1_506.dat", 20},             {"CN1220_ALPHA1_508.dat", 20},
    {"CL1220_ALPHA1_608.dat", 20},             {"CL1320_ALPHA1_606.dat", 20},
    {"CN9999_ALPHA1_brett.dat", 20},           {"CL9999_ALPHA1_brett.dat", 20},
    {"CM9999_ALPHA1_brett.dat", 20},           {"CX1420_ALPHA2_205.dat", 14},
    {"CY1620_ALPHA2_407.dat", 14},             {"CY1520_ALPHA2_409.dat", 14},
    {"CZ1420_ALPHA2_305.dat", 14},             {"CL1620_ALPHA2_607.dat", 14},
    {"CL1520_ALPHA2_609.dat", 14},             {"CM1420_ALPHA2_105.dat", 14},
    {"CN1620_ALPHA2_507.dat", 14},             {"CN1520_ALPHA2_509.dat", 14},
    {"ETA_DH1_brett.dat", 5}};

bool read_text(const std::string& path, int n, double* out, std::string& err) {
  FILE* f = fopen(path.c_str(), "r");
  if (!f) { err = "cannot open " + path; return false; }
  for (int i = 0; i < n; i++) {
    if (fscanf(f, "%lf", &out[i]) != 1) { fclose(f); err = "short read in " + path; return false; }
  }
  fclose(f);
  return true;
}

bool is_dir(const std::string& p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
bool is_file(const std::string& p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}
}  // namespace

size_t canon_table_offset(int t) {
  size_t off = F16_CANON_TABLES;
  for (int i = 0; i < t; i++) off += TABLES[i].n;
  return off;
}

static bool load_blob(const std::string& path, std::vector<double>& payload, std::string& err) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { err = "cannot open " + path; return false; }
  char magic[8];
  uint64_t n = 0;
  uint8_t sha[32], got[32];
  bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "F16AERO1", 8) == 0 && fread(&n, 8, 1, f) == 1 &&
            n == F16_CANON_DOUBLES && fread(sha, 1, 32, f) == 32;
  if (ok) {
    payload.resize(n);
    ok = fread(payload.data(), 8, n, f) == n;
  }
  fclose(f);
  if (!ok) { err = "malformed table blob " + path; return false; }
  sha_digest(payload.data(), payload.size() * 8, got);
  if (memcmp(sha, got, 32) != 0) { err = "sha256 mismatch in " + path; return false; }
  return true;
}

static bool load_dat_dir(const std::string& dir, std::vector<double>& payload, std::string& err) {
  payload.assign(F16_CANON_DOUBLES, 0.0);
  size_t off = 0;
  for (const TableFile& b : BREAKS) {
    if (!read_text(dir + "/" + b.file, b.n, payload.data() + off, err)) return false;
    off += b.n;
  }
  for (const TableFile& t : TABLES) {
    if (!read_text(dir + "/" + t.file, t.n, payload.data() + off, err)) return false;
    off += t.n;
  }
  return off == F16_CANON_DOUBLES;
}

bool load_canonical(const char* path, const std::string& lib_dir, std::vector<double>& payload, std::string& source,
                    std::string& err) {
  std::vector<std::string> cands;
  if (path && *path) {
    cands.push_back(path);
  } else {
    const char* env = getenv("F16_TABLE_PATH");
    if (env && *env) cands.push_back(env);
    if (!lib_dir.empty()) {
      cands.push_back(lib_dir + "/../data/f16_aero_v1.bin");
      cands.push_back(lib_dir + "/data/f16_aero_v1.bin");
      cands.push_back(lib_dir + "/f16_aero_v1.bin");
    }
    cands.push_back("./C");
    if (!lib_dir.empty()) cands.push_back(lib_dir);  // a shim living in the reference's own C/ directory
  }
  std::string last = "no candidate";
  for (const std::string& c : cands) {
    bool ok = false;
    if (is_dir(c)) {
      if (!is_file(c + "/ALPHA1.dat")) { last = "no ALPHA1.dat in " + c; continue; }
      ok = load_dat_dir(c, payload, last);
    } else if (is_file(c)) {
      ok = load_blob(c, payload, last);
    } else {
      last = "not found: " + c;
      continue;
    }
    if (ok) { source = c; return true; }
    if (path && *path) break;
  }
  err = "aero tables: " + last;
  return false;
}

void payload_sha256_hex(const std::vector<double>& payload, char out65[65]) {
  uint8_t d[32];
  sha_digest(payload.data(), payload.size() * 8, d);
  for (int i = 0; i < 32; i++) snprintf(out65 + 2 * i, 3, "%02x", d[i]);
}

// canonical (alpha fastest, 20 or 14 alpha points) -> device image (f16_tables.h)
void build_hifi_image(const std::vector<double>& p, bool clr_from_file, std::vector<double>& img) {
  img.assign(F16_IMG_HIFI_DOUBLES, 0.0);
  for (int i = 0; i < F16_IMG_NA; i++) img[F16_IMG_A + i] = p[F16_CANON_A1 + i];
  for (int i = 0; i < F16_N_B; i++) img[F16_IMG_B + i] = p[F16_CANON_B + i];
  for (int i = 0; i < F16_N_D1; i++) img[F16_IMG_D1 + i] = p[F16_CANON_D1 + i];
  for (int i = 0; i < F16_N_D2; i++) img[F16_IMG_D2 + i] = p[F16_CANON_D2 + i];
  auto tab = [&](int t) { return p.data() + canon_table_offset(t); };
  // value of table t at (ia, ib, id); na = its alpha count (20 for ALPHA1 grids, 14 for ALPHA2)
  auto at = [&](int t, int na, int ia, int ib, int id) { return tab(t)[(id * F16_N_B + ib) * na + ia]; };

  for (int i = 0; i < F16_N_D1; i++) img[F16_IMG_ETA + i] = tab(FT_eta_el)[i];

  const int g1_src[21] = {FT_CXq, FT_CYr, FT_CYp, FT_CZq, FT_CLr, FT_CLp, FT_CMq, FT_CNr, FT_CNp, FT_dCNbeta, FT_dCLbeta,
                          FT_dCm, FT_dCXq_lef, FT_dCYr_lef, FT_dCYp_lef, FT_dCZq_lef, FT_dCLr_lef, FT_dCLp_lef,
                          FT_dCMq_lef, FT_dCNr_lef, FT_dCNp_lef};
  for (int ia = 0; ia < F16_IMG_NA; ia++)
    for (int s = 0; s < 21; s++) {
      double v = tab(g1_src[s])[ia];
      // The reference never reads CL1320_ALPHA1_606.dat (hifi_F16_AeroData.c:965-972): as built, CLr == 0.
      if (g1_src[s] == FT_CLr && !clr_from_file) v = 0.0;
      img[F16_IMG_G1 + ia * F16_G1_STRIDE + s] = v;
    }

  for (int id = 0; id < F16_N_D2; id++)
    for (int ib = 0; ib < F16_N_B; ib++)
      for (int ia = 0; ia < F16_IMG_NA; ia++) {
        double* n = &img[F16_IMG_G3B + ((id * F16_N_B + ib) * F16_IMG_NA + ia) * F16_G3B_STRIDE];
        n[0] = at(FT_Cn, 20, ia, ib, id);
        n[1] = at(FT_Cl, 20, ia, ib, id);
      }
  for (int id = 0; id < F16_N_D1; id++)
    for (int ib = 0; ib < F16_N_B; ib++)
      for (int ia = 0; ia < F16_IMG_NA; ia++) {
        double* n = &img[F16_IMG_G3A + ((id * F16_N_B + ib) * F16_IMG_NA + ia) * F16_G3A_STRIDE];
        n[0] = at(FT_Cx, 20, ia, ib, id);
        n[1] = at(FT_Cz, 20, ia, ib, id);
        n[2] = at(FT_Cm, 20, ia, ib, id);
      }
  // alpha x beta group; dele = 0 is breakpoint 2 of DH1 and breakpoint 1 of DH2
  for (int ib = 0; ib < F16_N_B; ib++)
    for (int ia = 0; ia < F16_IMG_NA; ia++) {
      double* n = &img[F16_IMG_G2 + (ib * F16_IMG_NA + ia) * F16_G2_STRIDE];
      n[G2_Cx0] = at(FT_Cx, 20, ia, ib, 2);
      n[G2_Cz0] = at(FT_Cz, 20, ia, ib, 2);
      n[G2_Cm0] = at(FT_Cm, 20, ia, ib, 2);
      n[G2_Cy] = at(FT_Cy, 20, ia, ib, 0);
      n[G2_Cn0] = at(FT_Cn, 20, ia, ib, 1);
      n[G2_Cl0] = at(FT_Cl, 20, ia, ib, 1);
      n[G2_Cy_r30] = at(FT_Cy_r30, 20, ia, ib, 0);
      n[G2_Cn_r30] = at(FT_Cn_r30, 20, ia, ib, 0);
      n[G2_Cl_r30] = at(FT_Cl_r30, 20, ia, ib, 0);
      n[G2_Cy_a20] = at(FT_Cy_a20, 20, ia, ib, 0);
      n[G2_Cn_a20] = at(FT_Cn_a20, 20, ia, ib, 0);
      n[G2_Cl_a20] = at(FT_Cl_a20, 20, ia, ib, 0);
      n[G2_Cx_lef] = at(FT_Cx_lef, 14, ia, ib, 0);
      n[G2_Cz_lef] = at(FT_Cz_lef, 14, ia, ib, 0);
      n[G2_Cm_lef] = at(FT_Cm_lef, 14, ia, ib, 0);
      n[G2_Cy_lef] = at(FT_Cy_lef, 14, ia, ib, 0);
      n[G2_Cn_lef] = at(FT_Cn_lef, 14, ia, ib, 0);
      n[G2_Cl_lef] = at(FT_Cl_lef, 14, ia, ib, 0);
      n[G2_Cy_a20_lef] = at(FT_Cy_a20_lef, 14, ia, ib, 0);
      n[G2_Cn_a20_lef] = at(FT_Cn_a20_lef, 14, ia, ib, 0);
      n[G2_Cl_a20_lef] = at(FT_Cl_a20_lef, 14, ia, ib, 0);
    }
}

// canonical -> fast image (f16_fast.cuh): per alpha CELL the value at the lower node and the difference to the upper one
void build_hifi_fast_image(const std::vector<double>& p, bool clr_from_file, std::vector<double>& img) {
  img.assign(F16_FI_DOUBLES, 0.0);
  for (int i = 0; i < F16_IMG_NA; i++) img[F16_IMG_A + i] = p[F16_CANON_A1 + i];
  for (int i = 0; i < F16_N_B; i++) img[F16_IMG_B + i] = p[F16_CANON_B + i];
  for (int i = 0; i < F16_N_D1; i++) img[F16_IMG_D1 + i] = p[F16_CANON_D1 + i];
  for (int i = 0; i < F16_N_D2; i++) img[F16_IMG_D2 + i] = p[F16_CANON_D2 + i];
  auto tab = [&](int t) { return p.data() + canon_table_offset(t); };
  auto at = [&](int t, int na, int ia, int ib, int id) {
    // The reference never reads CL1320_ALPHA1_606.dat (hifi_F16_AeroData.c:965-972): as built, CLr == 0.
    if (t == FT_CLr && !clr_from_file) return 0.0;
    return tab(t)[(id * F16_N_B + ib) * na + ia];
  };
  auto put = [&](double* dst, int t, int na, int ia, int ib, int id) {
    const double f = at(t, na, ia, ib, id);
    dst[0] = f;
    dst[1] = at(t, na, ia + 1, ib, id) - f;
  };
  for (int i = 0; i < F16_N_D1 - 1; i++) {
    img[F16_FI_ETA + 2 * i] = tab(FT_eta_el)[i];
    img[F16_FI_ETA + 2 * i + 1] = tab(FT_eta_el)[i + 1] - tab(FT_eta_el)[i];
  }
  // rho = rho0 * tfac^4.14 (nlplant.c:478): centres c_i = (18.5 + i)/64; entry = (1/(64 c_i), 0.5 rho0 c_i^4.14)
  for (int i = 0; i < F16_FI_NPOW; i++) {
    const long double c = (18.5L + i) / 64.0L;
    img[F16_FI_POW + 2 * i] = (double)(1.0L / (64.0L * c));
    img[F16_FI_POW + 2 * i + 1] = (double)(0.5L * 2.377e-3L * powl(c, (long double)4.14));
  }
  // axis tables of fastmath::locate_hifi (check_grids() has verified that every breakpoint is a whole number of degrees)
  {
    double* ax = &img[F16_FI_AX];
    const double* A = &p[F16_CANON_A1];
    const double* B = &p[F16_CANON_B];
    const double* D1 = &p[F16_CANON_D1];
    for (int k = 0; k < F16_FI_NAC; k++) ax[F16_AX_A + k] = -A[k] / (A[k + 1] - A[k]);  // 4 - k
    for (int k = 0; k < F16_N_D1 - 1; k++) {
      ax[F16_AX_D1 + 2 * k] = 1.0 / (D1[k + 1] - D1[k]);
      ax[F16_AX_D1 + 2 * k + 1] = -D1[k] / (D1[k + 1] - D1[k]);
    }
    unsigned char lut[64] = {0};
    for (int t = 0; t <= 60; t++) {  // beta in [B[0] + t, B[0] + t + 1)
      int k = 0;
      while (k < F16_N_B - 2 && B[0] + t >= B[k + 1]) k++;
      lut[t] = (unsigned char)k;
      ax[F16_AX_TB + 2 * t] = 1.0 / (B[k + 1] - B[k]);
      ax[F16_AX_TB + 2 * t + 1] = -B[k] / (B[k + 1] - B[k]);
    }
    memcpy(ax + F16_AX_LUTB, lut, 64);
  }
  const int g1_src[FG1_COUNT][2] = {
      {FT_CXq, 20}, {FT_dCXq_lef, 14}, {FT_CZq, 20}, {FT_CMq, 20}, {FT_dCMq_lef, 14}, {FT_dCm, 20}, {FT_CYr, 20},
      {FT_dCYr_lef, 14}, {FT_CYp, 20}, {FT_dCYp_lef, 14}, {FT_CNr, 20}, {FT_dCNr_lef, 14}, {FT_CNp, 20}, {FT_dCNp_lef, 14},
      {FT_dCNbeta, 20}, {FT_CLr, 20}, {FT_dCLr_lef, 14}, {FT_CLp, 20}, {FT_dCLp_lef, 14}, {FT_dCLbeta, 20}};
  for (int ia = 0; ia < F16_FI_NAC; ia++)
    for (int s = 0; s < FG1_COUNT; s++) put(&img[F16_FI_G1 + ia * F16_FI_G1_STRIDE + 2 * s], g1_src[s][0], g1_src[s][1], ia, 0, 0);
  for (int id = 0; id < F16_N_D2; id++)
    for (int ib = 0; ib < F16_N_B; ib++)
      for (int ia = 0; ia < F16_FI_NAC; ia++) {
        double* n = &img[F16_FI_G3B + ((id * F16_N_B + ib) * F16_FI_NAC + ia) * F16_FI_G3B_STRIDE];
        put(n + 0, FT_Cn, 20, ia, ib, id);
        put(n + 2, FT_Cl, 20, ia, ib, id);
      }
  for (int id = 0; id < F16_N_D1; id++)
    for (int ib = 0; ib < F16_N_B; ib++)
      for (int ia = 0; ia < F16_FI_NAC; ia++) {
        double* n = &img[F16_FI_G3A + ((id * F16_N_B + ib) * F16_FI_NAC + ia) * F16_FI_G3A_STRIDE];
        put(n + 0, FT_Cx, 20, ia, ib, id);
        put(n + 2, FT_Cz, 20, ia, ib, id);
        put(n + 4, FT_Cm, 20, ia, ib, id);
      }
  // alpha x beta group: node values of the reference's delta coefficients (interpolation is linear, so the delta of
  // the interpolants is the interpolant of the node deltas); dele = 0 is breakpoint 2 of DH1 and breakpoint 1 of DH2
  auto node_val = [&](int slot, int ia, int ib) {
    auto A20 = [&](int base, int a20, int base_d) { return at(a20, 20, ia, ib, 0) - at(base, 20, ia, ib, base_d); };
    switch (slot) {
      case FG2_dCx_lef: return at(FT_Cx_lef, 14, ia, ib, 0) - at(FT_Cx, 20, ia, ib, 2);
      case FG2_dCz_lef: return at(FT_Cz_lef, 14, ia, ib, 0) - at(FT_Cz, 20, ia, ib, 2);
      case FG2_dCm_lef: return at(FT_Cm_lef, 14, ia, ib, 0) - at(FT_Cm, 20, ia, ib, 2);
      case FG2_Cy: return at(FT_Cy, 20, ia, ib, 0);
      case FG2_dCy_lef: return at(FT_Cy_lef, 14, ia, ib, 0) - at(FT_Cy, 20, ia, ib, 0);
      case FG2_dCy_a20: return A20(FT_Cy, FT_Cy_a20, 0);
      case FG2_dCy_a20_lef: return at(FT_Cy_a20_lef, 14, ia, ib, 0) - at(FT_Cy_lef, 14, ia, ib, 0) - A20(FT_Cy, FT_Cy_a20, 0);
      case FG2_dCy_r30: return at(FT_Cy_r30, 20, ia, ib, 0) - at(FT_Cy, 20, ia, ib, 0);
      case FG2_dCn_lef: return at(FT_Cn_lef, 14, ia, ib, 0) - at(FT_Cn, 20, ia, ib, 1);
      case FG2_dCn_a20: return A20(FT_Cn, FT_Cn_a20, 1);
      case FG2_dCn_a20_lef: return at(FT_Cn_a20_lef, 14, ia, ib, 0) - at(FT_Cn_lef, 14, ia, ib, 0) - A20(FT_Cn, FT_Cn_a20, 1);
      case FG2_dCn_r30: return at(FT_Cn_r30, 20, ia, ib, 0) - at(FT_Cn, 20, ia, ib, 1);
      case FG2_dCl_lef: return at(FT_Cl_lef, 14, ia, ib, 0) - at(FT_Cl, 20, ia, ib, 1);
      case FG2_dCl_a20: return A20(FT_Cl, FT_Cl_a20, 1);
      case FG2_dCl_a20_lef: return at(FT_Cl_a20_lef, 14, ia, ib, 0) - at(FT_Cl_lef, 14, ia, ib, 0) - A20(FT_Cl, FT_Cl_a20, 1);
      default: return at(FT_Cl_r30, 20, ia, ib, 0) - at(FT_Cl, 20, ia, ib, 1);  // FG2_dCl_r30
    }
  };
  for (int ib = 0; ib < F16_N_B; ib++)
    for (int ia = 0; ia < F16_FI_NAC; ia++)
      for (int s = 0; s < FG2_COUNT; s++) {
        double* dst = &img[F16_FI_G2 + (ib * F16_FI_NAC + ia) * F16_FI_G2_STRIDE + 2 * s];
        dst[0] = node_val(s, ia, ib);
        dst[1] = node_val(s, ia + 1, ib) - dst[0];
      }
}

static const double LOFI_DATA[F16_IMG_LOFI_DOUBLES] = {
#include "f16_lofi_data.inc"
};

void build_lofi_image(std::vector<double>& img) { img.assign(LOFI_DATA, LOFI_DATA + F16_IMG_LOFI_DOUBLES); }

bool check_grids(const std::vector<double>& p, std::string& err) {
  // the kernels' cell-index guesses assume the published grids; verify instead of trusting the files
  for (int i = 0; i < F16_N_A2; i++)
    if (p[F16_CANON_A1 + i] != p[F16_CANON_A2 + i] || p[F16_CANON_A1 + i] != -20.0 + 5.0 * i) { err = "unexpected ALPHA grid"; return false; }
  const double B[19] = {-30, -25, -20, -15, -10, -8, -6, -4, -2, 0, 2, 4, 6, 8, 10, 15, 20, 25, 30};
  for (int i = 0; i < 19; i++) if (p[F16_CANON_B + i] != B[i]) { err = "unexpected BETA1 grid"; return false; }
  const double D1[5] = {-25, -10, 0, 10, 25}, D2[3] = {-25, 0, 25};
  for (int i = 0; i < 5; i++) if (p[F16_CANON_D1 + i] != D1[i]) { err = "unexpected DH1 grid"; return false; }
  for (int i = 0; i < 3; i++) if (p[F16_CANON_D2 + i] != D2[i]) { err = "unexpected DH2 grid"; return false; }
  return true;
}

}  // namespace f16
