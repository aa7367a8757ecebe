// =============================================================================
// qilqr_riccati_step.cuh -- one knot of the backward Riccati recursion for a quad of lanes
// (ilqr.hh:118-140).  Shared by the fused kernel (k_backward_g4: records in per-problem shared
// memory, stride RS = 1) and the split kernel (k_riccati_g4: record tiles [elem][8 problems]
// brought in by TMA bulk copies, stride RS = 8).  See qilqr_backward_g4.cuh for the algorithm.
//   rec    : this knot's linearisation record, element e at rec[e * RS]
//   gk_lane, gK_lane : where this lane's gains go -- gk + c * B + b and gK + 3 c * B + b of the problem (the padding
//            quads of a last tile shadow the last problem and rewrite identical values; quads of the persistent
//            tail kernel whose problem is not in its backward pass get a scratch array)
//   s2Qvv  : 2*Q_vv (6x6): dense for RS = 1; for RS = 8 the quad's slot of the tile-like table (g4::QVV_TILE)
//   xch    : this problem's exchange area (g4::XCH doubles)
//   V0..V3 : the lane's column block c of V_xx (in/out);  vx: v_x (replicated, in/out)
//   V88    : V_xx[8:12, 8:12] (replicated, in/out) -- all that Q_uu = C_uu + B^T V_xx B needs
// =============================================================================
#pragma once
// (included by qilqr_backward_g4.cuh after the g4 helpers it uses)

namespace qilqr {
namespace g4 {
template <int RS>
QD void ld9s(const double *s, double *r) {
#pragma unroll
  for (int e = 0; e < 9; ++e) r[e] = s[e * RS];
}

struct NoRecordHook {
  static constexpr bool kActive = false;
  QD void operator()() const {}
};
// record_done(): called once the whole warp has made its last read of this knot's record (a kernel with a single
// record buffer starts the next knot's copy there).
template <int RS, bool DENSEQ = false, class RecordDone = NoRecordHook>
QD void riccati_step(const DeviceParams &p, const BackwardArgs &a, const double *rec, const double *s2Qvv,
                     double *xch, const int c, double *gk_lane, double *gK_lane, const int ii, const int B, double *V0,
                     double *V1, double *V2, double *V3, double *vx, double *V88, double &QuTk, double &kTQuuk,
                     RecordDone record_done = RecordDone()) {
  const double dgz[3] = {rec[R_GZ * RS], rec[(R_GZ + 1) * RS], rec[(R_GZ + 2) * RS]};
  const double ndgz[3] = {-dgz[0], -dgz[1], -dgz[2]};

  // ---------------- step 1: M[:,c] = A^T V[:,c] ----------------
  {
    double Ab[9], Mb[9];
    ld9s<RS>(rec + R_RE * RS, Ab);
    m3_mulT(Ab, V0, Mb);
    st9(xch + moff(0, c), Mb);
    double Tb[9];
    ld9s<RS>(rec + R_TE * RS, Tb);
    m3_mulT(Tb, V0, Mb);
    m3_maddT(Ab, V1, Mb);
    m3_hat_madd(ndgz, V2, Mb);  // dG^T = hat(dgz)^T = hat(-dgz)
    st9(xch + moff(1, c), Mb);
    ld9s<RS>(rec + R_DJR * RS, Ab);
    m3_mulT(Ab, V0, Mb);
#pragma unroll
    for (int e = 0; e < 9; ++e) Mb[e] += V2[e];
    st9(xch + moff(2, c), Mb);
    ld9s<RS>(rec + R_DQB * RS, Tb);
    m3_mulT(Tb, V0, Mb);
    m3_maddT(Ab, V1, Mb);
    ld9s<RS>(rec + R_WD * RS, Tb);
    m3_maddT(Tb, V3, Mb);
    st9(xch + moff(3, c), Mb);
  }

  // ---------------- step 2 (replicated): Q_uu, Q_u, Q_x, factorisation, k ----------------
  // (independent of step 1: no barrier in between, so the factorisation's dependency chain overlaps the
  //  block products above)
  double Quu[16], Qu[4], Qx[12], k[4];
  Ldlt4 f;
  {
    double BtV[16];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        double acc = p.Bu[jj] * V88[cc];
#pragma unroll
        for (int r = 1; r < 4; ++r) acc = QFMA(p.Bu[4 * r + jj], V88[4 * r + cc], acc);
        BtV[4 * jj + cc] = acc;
      }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        double acc = BtV[4 * jj] * p.Bu[l];
#pragma unroll
        for (int cc = 1; cc < 4; ++cc) acc = QFMA(BtV[4 * jj + cc], p.Bu[4 * cc + l], acc);
        Quu[4 * jj + l] = QFMA(2.0, p.R[4 * jj + l], acc);
      }
    if (p.quu_reg != 0.0) {
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) Quu[5 * jj] += p.quu_reg;
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      double acc = p.Bu[jj] * vx[8];
#pragma unroll
      for (int r = 1; r < 4; ++r) acc = QFMA(p.Bu[4 * r + jj], vx[8 + r], acc);
      Qu[jj] = rec[(R_CU + jj) * RS] + acc;
    }
    // Q.x = C.x + A^T v_x
    double Ab[9], T[12];
    ld9s<RS>(rec + R_RE * RS, Ab);
    m3T_vec(Ab, vx, T);
    double Tb[9];
    ld9s<RS>(rec + R_TE * RS, Tb);
    m3T_vec(Tb, vx, T + 3);
    m3T_vec_add(Ab, vx + 3, T + 3);
    {  // += dG^T vx[6:9] = hat(-dgz) vx[6:9]
      const double *v = vx + 6;
      T[3] = QFMA(ndgz[1], v[2], QFMA(-ndgz[2], v[1], T[3]));
      T[4] = QFMA(-ndgz[0], v[2], QFMA(ndgz[2], v[0], T[4]));
      T[5] = QFMA(ndgz[0], v[1], QFMA(-ndgz[1], v[0], T[5]));
    }
    ld9s<RS>(rec + R_DJR * RS, Ab);
    m3T_vec(Ab, vx, T + 6);
#pragma unroll
    for (int e = 0; e < 3; ++e) T[6 + e] += vx[6 + e];
    ld9s<RS>(rec + R_DQB * RS, Tb);
    m3T_vec(Tb, vx, T + 9);
    m3T_vec_add(Ab, vx + 3, T + 9);
    ld9s<RS>(rec + R_WD * RS, Tb);
    m3T_vec_add(Tb, vx + 9, T + 9);
#pragma unroll
    for (int e = 0; e < 12; ++e) Qx[e] = rec[(R_CX + e) * RS] + T[e];
#pragma unroll
    for (int e = 0; e < 16; ++e) f.m[e] = Quu[e];
    ldlt4_compute(f);
    double rhs[4] = {Qu[0], Qu[1], Qu[2], Qu[3]};
    ldlt4_solve(f, rhs);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) k[jj] = -rhs[jj];
  }

  __syncwarp();

  // ---------------- step 3: row block r = c of Q_xx and Q_xu ----------------
  double Q0[9], Q1[9], Q2[9], Q3[9], Qxu[12];
  {
    // C_xx[r,:] (cost.hh:52): rows 0..5 = [C_pp | C_pv] from the record, rows 6..11 = [C_vp | 2 Q_vv];
    // C_pv and C_vp vanish (and are not stored) when Q has no pose/velocity coupling (!DENSEQ).
    const bool lo = c < 2;
    if (DENSEQ) {
      const double *srcA = lo ? (rec + (R_CPP + R_CPP_GROUP * c) * RS) : (rec + (R_CVP + 18 * (c - 2)) * RS);
      const double *srcB = lo ? (rec + (R_CPV + 18 * c) * RS) : (s2Qvv + qvv_group<RS>(c - 2));
      constexpr int sB = RS;  // record and table have the same element stride
#pragma unroll
      for (int ri = 0; ri < 3; ++ri)
#pragma unroll
        for (int cj = 0; cj < 3; ++cj) {
          Q0[3 * ri + cj] = srcA[(6 * ri + cj) * RS];
          Q1[3 * ri + cj] = srcA[(6 * ri + 3 + cj) * RS];
          Q2[3 * ri + cj] = srcB[(6 * ri + cj) * sB];
          Q3[3 * ri + cj] = srcB[(6 * ri + 3 + cj) * sB];
        }
    } else {
      // Lanes 0, 1 own pose rows [C_pp | 0], lanes 2, 3 velocity rows [0 | 2 Q_vv]: one 3x6 source per lane, and the
      // zero half enters as source * 0.0 inside the fused add below instead of a select per element.
      const double *src = lo ? (rec + (R_CPP + R_CPP_GROUP * c) * RS) : (s2Qvv + qvv_group<RS>(c - 2));
      constexpr int sst = RS;  // record and table have the same element stride
#pragma unroll
      for (int ri = 0; ri < 3; ++ri)
#pragma unroll
        for (int cj = 0; cj < 3; ++cj) {
          Q0[3 * ri + cj] = src[(6 * ri + cj) * sst];
          Q1[3 * ri + cj] = src[(6 * ri + 3 + cj) * sst];
        }
    }
    // C_xx[r,:] + (A^T V A)[r,:].  cadd(C, m, T) = C * m + T with m = 1 (exact: C + T) or 0 (T, up to the sign of a
    // zero T); DENSEQ adds the stored blocks directly.
    const double mlo = lo ? 1.0 : 0.0, mhi = lo ? 0.0 : 1.0;
    double X0[9], X1[9], X2[9], X3[9], Ab[9], T[9];
    ld9(xch + moff(c, 0), X0);
    ld9s<RS>(rec + R_RE * RS, Ab);
    m3_mul(X0, Ab, T);
    double C0[9], C1[9];  // the lane's 3x6 source (!DENSEQ)
#pragma unroll
    for (int e = 0; e < 9; ++e) {
      C0[e] = Q0[e];
      C1[e] = Q1[e];
      Q0[e] = DENSEQ ? Q0[e] + T[e] : QFMA(C0[e], mlo, T[e]);
    }
    double Tb[9];
    ld9s<RS>(rec + R_TE * RS, Tb);
    m3_mul(X0, Tb, T);
    ld9(xch + moff(c, 1), X1);
    m3_madd(X1, Ab, T);
    ld9(xch + moff(c, 2), X2);
    m3_madd_hat(X2, dgz, T);
#pragma unroll
    for (int e = 0; e < 9; ++e) Q1[e] = DENSEQ ? Q1[e] + T[e] : QFMA(C1[e], mlo, T[e]);
    ld9s<RS>(rec + R_DJR * RS, Ab);
    m3_mul(X0, Ab, T);
#pragma unroll
    for (int e = 0; e < 9; ++e) Q2[e] = DENSEQ ? Q2[e] + (T[e] + X2[e]) : QFMA(C0[e], mhi, T[e] + X2[e]);
    ld9s<RS>(rec + R_DQB * RS, Tb);
    m3_mul(X0, Tb, T);
    m3_madd(X1, Ab, T);
    ld9(xch + moff(c, 3), X3);
    ld9s<RS>(rec + R_WD * RS, Tb);
    m3_madd(X3, Tb, T);
#pragma unroll
    for (int e = 0; e < 9; ++e) Q3[e] = DENSEQ ? Q3[e] + T[e] : QFMA(C1[e], mhi, T[e]);
    // Q.xu[r] = M[r, 8] B[8,:] + M[r, 9:12] B[9:12,:]
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        double acc = X2[3 * s + 2] * p.Bu[jj];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) acc = QFMA(X3[3 * s + cc], p.Bu[4 * (1 + cc) + jj], acc);
        Qxu[4 * s + jj] = acc;
      }
  }

  if (RecordDone::kActive) {
    __syncwarp();
    record_done();
  }

  // ---------------- step 4: K[:, 3r..3r+2], (K^T Q_uu)[r], v_x'[r] ----------------
  double KQ[12];  // [s][l]
  {
    double Ks[12];  // [j][s]
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      double rhs[4] = {Qxu[4 * s], Qxu[4 * s + 1], Qxu[4 * s + 2], Qxu[4 * s + 3]};
      ldlt4_solve(f, rhs);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) Ks[3 * jj + s] = -rhs[jj];
    }
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        double acc = Ks[s] * Quu[l];
#pragma unroll
        for (int jj = 1; jj < 4; ++jj) acc = QFMA(Ks[3 * jj + s], Quu[4 * jj + l], acc);
        KQ[4 * s + l] = acc;
      }
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      double acc = KQ[4 * s] * k[0];
#pragma unroll
      for (int l = 1; l < 4; ++l) acc = QFMA(KQ[4 * s + l], k[l], acc);
      const double qx = sel4(c, Qx[s], Qx[3 + s], Qx[6 + s], Qx[9 + s]);
      xch[X_VX + 3 * c + s] = qx - acc;
    }
    // rows 12 jj + 3 c + s of K and row c of k at this knot: one 64-bit base per knot, 32-bit row offsets
    double *gKk = gK_lane + size_t(ii) * 48 * size_t(B);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        xch[X_K + 12 * jj + 3 * c + s] = Ks[3 * jj + s];
        gKk[(12 * jj + s) * B] = Ks[3 * jj + s];
      }
    gk_lane[size_t(ii) * 4 * size_t(B)] = sel4(c, k[0], k[1], k[2], k[3]);
  }
  __syncwarp();

  // ---------------- step 5: V'[r,:] = Q_xx[r,:] - (K^T Q_uu)[r] K ----------------
  {
    double Kall[48];
#pragma unroll
    for (int e = 0; e < 48; ++e) Kall[e] = xch[X_K + e];
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        double a0 = KQ[4 * s] * Kall[cc], a1 = KQ[4 * s] * Kall[3 + cc], a2 = KQ[4 * s] * Kall[6 + cc],
               a3 = KQ[4 * s] * Kall[9 + cc];
#pragma unroll
        for (int l = 1; l < 4; ++l) {
          a0 = QFMA(KQ[4 * s + l], Kall[12 * l + cc], a0);
          a1 = QFMA(KQ[4 * s + l], Kall[12 * l + 3 + cc], a1);
          a2 = QFMA(KQ[4 * s + l], Kall[12 * l + 6 + cc], a2);
          a3 = QFMA(KQ[4 * s + l], Kall[12 * l + 9 + cc], a3);
        }
        Q0[3 * s + cc] -= a0;
        Q1[3 * s + cc] -= a1;
        Q2[3 * s + cc] -= a2;
        Q3[3 * s + cc] -= a3;
      }
#pragma unroll
    for (int e = 0; e < 12; ++e) vx[e] = xch[X_VX + e];
    st9(xch + moff(c, 0), Q0);
    st9(xch + moff(c, 1), Q1);
    st9(xch + moff(c, 2), Q2);
    st9(xch + moff(c, 3), Q3);
    // expected cost reduction terms (ilqr.hh:136-140), replicated
    double acc = Qu[0] * k[0];
#pragma unroll
    for (int jj = 1; jj < 4; ++jj) acc = QFMA(Qu[jj], k[jj], acc);
    QuTk = QuTk + acc;
    double acc2 = 0.0;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      double z = k[0] * Quu[l];
#pragma unroll
      for (int jj = 1; jj < 4; ++jj) z = QFMA(k[jj], Quu[4 * jj + l], z);
      acc2 = (l == 0) ? z * k[0] : QFMA(z, k[l], acc2);
    }
    kTQuuk = kTQuuk + acc2;
  }
  __syncwarp();

  // ---------------- step 6: V[:,c] <- V'[:,c] ----------------
  ld9(xch + moff(0, c), V0);
  ld9(xch + moff(1, c), V1);
  ld9(xch + moff(2, c), V2);
  ld9(xch + moff(3, c), V3);
  V88[0] = xch[moff(2, 2) + 8];
#pragma unroll
  for (int e = 0; e < 3; ++e) {
    V88[1 + e] = xch[moff(2, 3) + 6 + e];
    V88[4 * (1 + e)] = xch[moff(3, 2) + 3 * e + 2];
    V88[5 + e] = xch[moff(3, 3) + e];
    V88[9 + e] = xch[moff(3, 3) + 3 + e];
    V88[13 + e] = xch[moff(3, 3) + 6 + e];
  }
  if (p.symmetrize_vxx) {  // V <- (V + V^T)/2: block (K,c) needs (block (c,K))^T, which this lane owns
#pragma unroll
    for (int ri = 0; ri < 3; ++ri)
#pragma unroll
      for (int cj = 0; cj < 3; ++cj) {
        V0[3 * ri + cj] = 0.5 * (V0[3 * ri + cj] + Q0[3 * cj + ri]);
        V1[3 * ri + cj] = 0.5 * (V1[3 * ri + cj] + Q1[3 * cj + ri]);
        V2[3 * ri + cj] = 0.5 * (V2[3 * ri + cj] + Q2[3 * cj + ri]);
        V3[3 * ri + cj] = 0.5 * (V3[3 * ri + cj] + Q3[3 * cj + ri]);
      }
#pragma unroll
    for (int r8 = 0; r8 < 4; ++r8)
#pragma unroll
      for (int c8 = r8 + 1; c8 < 4; ++c8) {
        const double m8 = 0.5 * (V88[4 * r8 + c8] + V88[4 * c8 + r8]);
        V88[4 * r8 + c8] = m8;
        V88[4 * c8 + r8] = m8;
      }
  }
  __syncwarp();
}
}  // namespace g4
}  // namespace qilqr
