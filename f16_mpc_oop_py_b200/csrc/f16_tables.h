// f16_tables.h -- layouts of the aero database: the canonical payload (what tools/pack_tables.py writes, the
// reference's own per-file column-major order) and the device image the kernels gather from.
//
// Reference: the 43 accessors of C/hifi_F16_AeroData.c:109-1861 each own one table on a grid of
// ALPHA1(20) / ALPHA2(14) x BETA1(19) x DH1(5) / DH2(3) (breakpoint loaders :7-105); the lofi tables are
// the array initialisers of C/lofi_F16_AeroData.c:17-26,66-104,192-206,271-283,343-344.
#pragma once

// ---------------------------------------------------------------------------------------------------
// canonical payload (doubles): breakpoints then tables, in this order
// ---------------------------------------------------------------------------------------------------
#define F16_N_A1 20
#define F16_N_A2 14
#define F16_N_B 19
#define F16_N_D1 5
#define F16_N_D2 3

#define F16_CANON_A1 0
#define F16_CANON_A2 20
#define F16_CANON_B 34
#define F16_CANON_D1 53
#define F16_CANON_D2 58
#define F16_CANON_TABLES 61
#define F16_CANON_DOUBLES 13466

// canonical table ids (order of HIFI_TABLES in tools/pack_tables.py)
enum F16CanonTable {
  FT_Cx = 0, FT_Cz, FT_Cm,                                                        // A1 x B x D1 (1900)
  FT_Cn, FT_Cl,                                                                   // A1 x B x D2 (1140)
  FT_Cy, FT_Cy_r30, FT_Cn_r30, FT_Cl_r30, FT_Cy_a20, FT_Cn_a20, FT_Cl_a20,        // A1 x B (380)
  FT_Cx_lef, FT_Cz_lef, FT_Cm_lef, FT_Cy_lef, FT_Cn_lef, FT_Cl_lef,
  FT_Cy_a20_lef, FT_Cn_a20_lef, FT_Cl_a20_lef,                                    // A2 x B (266)
  FT_CXq, FT_CZq, FT_CMq, FT_CYp, FT_CYr, FT_CNr, FT_CNp, FT_CLp, FT_CLr,
  FT_dCNbeta, FT_dCLbeta, FT_dCm,                                                 // A1 (20)
  FT_dCXq_lef, FT_dCYr_lef, FT_dCYp_lef, FT_dCZq_lef, FT_dCLr_lef, FT_dCLp_lef,
  FT_dCMq_lef, FT_dCNr_lef, FT_dCNp_lef,                                          // A2 (14)
  FT_eta_el,                                                                      // D1 (5)
  FT_COUNT
};

// ---------------------------------------------------------------------------------------------------
// device image, hifi (doubles).  Only alpha <= 45 deg is kept (14 alpha points): the *_lef tables stop at
// 45 deg, beyond which the reference's Nlplant is undefined (mexndinterp.c:121-123), and ALPHA2 equals
// ALPHA1[0:14], so one alpha cell serves every table.  Tables that share a grid are interleaved so that
// one grid node is one contiguous, 16-byte aligned run (node-major, alpha fastest among nodes).
// ---------------------------------------------------------------------------------------------------
#define F16_IMG_NA 14
#define F16_IMG_A 0        // 14 alpha breakpoints (+2 pad)
#define F16_IMG_B 16       // 19 beta breakpoints (+1 pad)
#define F16_IMG_D1 36      // 5 (+1 pad)
#define F16_IMG_D2 42      // 3 (+3 pad)
#define F16_IMG_ETA 48     // eta_el on DH1: 5 (+3 pad)
#define F16_IMG_G1 56      // alpha-only group: 14 nodes x 22
#define F16_G1_STRIDE 22
#define F16_IMG_G3B 364    // (Cn, Cl) on alpha x beta x DH2: 14*19*3 nodes x 2
#define F16_G3B_STRIDE 2
#define F16_IMG_G3A 1960   // (Cx, Cz, Cm, pad) on alpha x beta x DH1: 14*19*5 nodes x 4
#define F16_G3A_STRIDE 4
#define F16_IMG_G2 7280    // alpha x beta group: 14*19 nodes x 22
#define F16_G2_STRIDE 22
#define F16_IMG_HIFI_DOUBLES 13132
#define F16_IMG_HIFI_BYTES (F16_IMG_HIFI_DOUBLES * 8)

// slots inside a G1 node (functions of alpha only)
enum F16G1Slot {
  G1_CXq = 0, G1_CYr, G1_CYp, G1_CZq, G1_CLr, G1_CLp, G1_CMq, G1_CNr, G1_CNp,
  G1_dCNbeta, G1_dCLbeta, G1_dCm,
  G1_dCXq_lef, G1_dCYr_lef, G1_dCYp_lef, G1_dCZq_lef, G1_dCLr_lef, G1_dCLp_lef, G1_dCMq_lef, G1_dCNr_lef,
  G1_dCNp_lef, G1_PAD
};

// slots inside a G2 node (functions of alpha, beta); *0 = the dele = 0 slice of the 3-D table
enum F16G2Slot {
  G2_Cx0 = 0, G2_Cz0, G2_Cm0, G2_Cy, G2_Cn0, G2_Cl0,
  G2_Cy_r30, G2_Cn_r30, G2_Cl_r30, G2_Cy_a20, G2_Cn_a20, G2_Cl_a20,
  G2_Cx_lef, G2_Cz_lef, G2_Cm_lef, G2_Cy_lef, G2_Cn_lef, G2_Cl_lef,
  G2_Cy_a20_lef, G2_Cn_a20_lef, G2_Cl_a20_lef, G2_PAD
};

// ---------------------------------------------------------------------------------------------------
// device image, hifi, "fast" layout (f16_fast.cuh): per alpha CELL the value at the lower node f and the difference
// d to the upper node, interleaved (f, d), so one LDS.128 + one FMA is an alpha interpolation.
// ---------------------------------------------------------------------------------------------------
// the first 48 doubles (breakpoints) equal the strict image so hifi_locate() serves both
#define F16_FI_NAC 13        // alpha cells (-20:5:45)
#define F16_FI_ETA 48        // eta_el per DH1 cell: (f, d) x 4
#define F16_FI_POW 56        // 48 x (1/(64 c_i), 0.5*rho0*c_i^4.14), c_i = (18.5 + i)/64
#define F16_FI_NPOW 48
#define F16_FI_G1 152        // alpha-only group: 13 cells x 20 tables x (f, d) (+2 pad: cell stride 336 B = 80 mod 128)
#define F16_FI_G1_STRIDE 42
#define F16_FI_G3B 698       // (Cn, Cl) on DH2 x beta x alpha-cell: 3*19*13 nodes x 2 x (f, d)
#define F16_FI_G3B_STRIDE 4
#define F16_FI_G3A 3662      // (Cx, Cz, Cm) on DH1 x beta x alpha-cell: 5*19*13 nodes x 3 x (f, d)
#define F16_FI_G3A_STRIDE 6
#define F16_FI_G2 11072      // alpha x beta group: 19*13 nodes x 16 tables x (f, d) (+2 pad: node stride 272 B = 16 mod 128,
#define F16_FI_G2_STRIDE 34    //   so that lanes in neighbouring cells do not meet in the same shared-memory banks)
// axis tables of locate_hifi(): the cell of a query is found from the INTEGER part of (query - axis start) -- every breakpoint of
// ALPHA1 / BETA1 / DH1 / DH2 is a whole number of degrees -- and the weight is one fma(query, 1 / cell width, offset of the cell)
#define F16_FI_AX 19470      // start of the axis tables (16-byte aligned)
#define F16_AX_A 0           // alpha: 4 - k for cell k = 0..12 (+3 pad): la = fma(alpha, 0.2, 4 - k)
#define F16_AX_D1 16         // DH1: (1 / width, -(lower breakpoint) / width) for cell 0..3
#define F16_AX_LUTB 24       // BETA1: 64 bytes, cell of floor(beta + 30) = 0..60
#define F16_AX_TB 32         // BETA1: (1 / width, -(lower breakpoint) / width) of that cell, again by floor(beta + 30): 61 x 2
#define F16_FI_DOUBLES 19624
#define F16_FI_BYTES (F16_FI_DOUBLES * 8)

enum F16FastG1 {  // table order inside a G1 cell
  FG1_Cxq = 0, FG1_dCxq_lef, FG1_Czq, FG1_Cmq, FG1_dCmq_lef, FG1_dCm, FG1_Cyr, FG1_dCyr_lef, FG1_Cyp, FG1_dCyp_lef,
  FG1_Cnr, FG1_dCnr_lef, FG1_Cnp, FG1_dCnp_lef, FG1_dCnbeta, FG1_Clr, FG1_dClr_lef, FG1_Clp, FG1_dClp_lef, FG1_dClbeta,
  FG1_COUNT
};
enum F16FastG2 {  // table order inside a G2 node: the delta coefficients of hifi_C_lef / hifi_rudder / hifi_ailerons
  FG2_dCx_lef = 0, FG2_dCz_lef, FG2_dCm_lef,                       // C_lef(a,b) - C(a,b,0)            hifi:1892-1899
  FG2_Cy, FG2_dCy_lef, FG2_dCy_a20, FG2_dCy_a20_lef, FG2_dCy_r30,  // Cy; C_a20 - C; C_a20_lef - C_lef - (C_a20 - C); C_r30 - C
  FG2_dCn_lef, FG2_dCn_a20, FG2_dCn_a20_lef, FG2_dCn_r30,          //                                   hifi:1913-1926
  FG2_dCl_lef, FG2_dCl_a20, FG2_dCl_a20_lef, FG2_dCl_r30,
  FG2_COUNT
};

// ---------------------------------------------------------------------------------------------------
// lofi image (doubles), exactly the flat array of f16_lofi_data.inc
// ---------------------------------------------------------------------------------------------------
#define F16_LOFI_DAMP 0      // [9][12]
#define F16_LOFI_DMOM 108    // [4: ALA ALR ANA ANR][8 rows: 7 + one zero row][12]
#define F16_LOFI_CLCN 492    // [2: AL AN][7][12]
#define F16_LOFI_CXCM 660    // [2: AX AM][5][12]
#define F16_LOFI_CZ 780      // [12]
#define F16_IMG_LOFI_DOUBLES 792
#define F16_IMG_LOFI_BYTES (F16_IMG_LOFI_DOUBLES * 8)
