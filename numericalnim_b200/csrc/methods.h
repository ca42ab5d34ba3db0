// methods.h — the integrator table of the product (its own copy of the Butcher data; the test-suite
// cross-checks it bit for bit against tests/golden/tableaux.json, which was extracted mechanically from
// the reference source, and against the CPU oracle).
//
// Every row is the literal term list of the corresponding expression in numericalnim's ode.nim, in
// source order, INCLUDING the zero weights the reference multiplies through (a72, bHat2, Vern65's
// zeros). The launcher drops zero-weight terms (saves a read; bitwise identical for finite data)
// unless the context is in strict_zeros mode.
#pragma once
#include <cstring>
#include <initializer_list>
#include <utility>

namespace b200rk {

constexpr int kMaxStages = 9;

struct Row {
  int m = 0;
  int idx[kMaxStages] = {0};   // 1-based stage index of each term
  double w[kMaxStages] = {0};
};

struct MethodDef {
  const char* name = "";
  int stages = 0;
  bool use_fsal = false;       // as passed to ODESolver (ode.nim:609-649)
  double order = 0;            // float `order` of the outer controller (ode.nim:474,537)
  int order_int = 0;           // int `order` of commonAdaptiveMethodCode (ode.nim:57,71)
  bool adaptive = false;
  bool k1_from_fsal = false;   // k1 = FSAL (pairs) vs k1 = f(t, y) on every attempt
  double c[kMaxStages + 1] = {0};     // stage time t + dt*c[s], s = 2..stages
  Row a[kMaxStages + 1];              // stage-s input row
  double a_cfac[kMaxStages + 1] = {0}; // scalar in front of the bracket: (cfac*dt), cfac == 1 -> dt
  bool a_chain[kMaxStages + 1] = {false}; // Kutta3 stage 3: ((y + (w0*dt)*k) + (w1*dt)*k)
  Row b;  double b_cfac = 1.0;        // yNew = y + (b_cfac*dt)*(b . k)
  bool rk4_final = false;             // y + dt/6*(k1 + 2*(k2+k3) + k4)
  Row bhat; double bhat_cfac = 1.0;   // yLow row (or the direct error row)
  bool err_direct = false;            // Tsit54: error_y = dt*(bHat . k)
  bool ynew_is_last_stage_input = false; // stage-S input row == b row literally (DOPRI54, Tsit54, BS32)
  int fsal_out = 0;                   // stage whose derivative is returned as FSAL (0: returns yNew)
};

inline Row make_row(std::initializer_list<std::pair<int, double>> terms) {
  Row r;
  for (auto& t : terms) { r.idx[r.m] = t.first; r.w[r.m] = t.second; ++r.m; }
  return r;
}
inline Row dense_row(const double* w, int m) {
  Row r; r.m = m;
  for (int j = 0; j < m; ++j) { r.idx[j] = j + 1; r.w[j] = w[j]; }
  return r;
}

// ---- the three FSAL pairs (ode.nim:237-468) ---------------------------------------------------------
inline MethodDef make_fsal_pair(const char* name, int stages, int order, const double* c /*[stages+1]*/,
                           const double (*a)[kMaxStages], const double* b, int nb, const double* bhat, int nbh,
                           bool direct, bool ynew_last) {
  MethodDef m;
  m.name = name; m.stages = stages; m.use_fsal = true; m.order = double(order); m.order_int = order;
  m.adaptive = true; m.k1_from_fsal = true;
  for (int s = 2; s <= stages; ++s) { m.c[s] = c[s]; m.a[s] = dense_row(a[s], s - 1); m.a_cfac[s] = 1.0; }
  m.b = dense_row(b, nb); m.bhat = dense_row(bhat, nbh);
  m.err_direct = direct; m.ynew_is_last_stage_input = ynew_last; m.fsal_out = stages;
  return m;
}

inline MethodDef make_dopri54() {  // ode.nim:240-282
  static const double c[8] = {0, 0, 1.0 / 5.0, 3.0 / 10.0, 4.0 / 5.0, 8.0 / 9.0, 1.0, 1.0};
  static const double a[8][kMaxStages] = {
      {0}, {0},
      {1.0 / 5.0},
      {3.0 / 40.0, 9.0 / 40.0},
      {44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0},
      {19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0},
      {9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0},
      {35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0}};
  static const double bhat[7] = {5179.0 / 57600.0, 0.0, 7571.0 / 16695.0, 393.0 / 640.0,
                                 -92097.0 / 339200.0, 187.0 / 2100.0, 1.0 / 40.0};
  return make_fsal_pair("dopri54", 7, 5, c, a, a[7], 6, bhat, 7, false, true);  // b_i = a7i (ode.nim:269-274)
}

inline MethodDef make_tsit54() {  // ode.nim:310-352
  static const double c[8] = {0, 0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0};
  static const double a[8][kMaxStages] = {
      {0}, {0},
      {0.161},
      {-0.008480655492356989, 0.335480655492357},
      {2.8971530571054935, -6.359448489975075, 4.3622954328695815},
      {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525},
      {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383},
      {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}};
  static const double bhat[7] = {-0.001780011052226, -0.000816434459657, 0.007880878010262, -0.144711007173263,
                                 0.582357165452555, -0.458082105929187, 1.0 / 66.0};
  return make_fsal_pair("tsit54", 7, 5, c, a, a[7], 6, bhat, 7, true, true);  // b_i = a7i (ode.nim:339-344)
}

inline MethodDef make_vern65() {  // ode.nim:380-443
  static const double c[10] = {0, 0, 0.06, 0.09593333333333333, 0.1439, 0.4973, 0.9725, 0.9995, 1.0, 1.0};
  static const double a[10][kMaxStages] = {
      {0}, {0},
      {0.06},
      {0.019239962962962962, 0.07669337037037037},
      {0.035975, 0.0, 0.107925},
      {1.3186834152331484, 0.0, -5.042058063628562, 4.220674648395414},
      {-41.87259166432751, 0.0, 159.43256216313748, -122.11921356501004, 5.531743066200053},
      {-54.430156935316504, 0.0, 207.06725136501848, -158.61081378459, 6.991816585950242, -0.01859723106220323},
      {-54.66374178728198, 0.0, 207.95280625538936, -159.2889574744995, 7.018743740796944, -0.018338785905045722,
       -0.0005119484997882099},
      {0.03438957868357036, 0.0, 0.0, 0.25826245556335037, 0.4209371189673537, 4.405396469669310,
       -176.48311902429865, 172.36413340141507}};
  // The sixth-order weights are their own literals (ode.nim:426-433); b4 and b7 differ from a94/a97 by
  // one ulp, so yNew is NOT the stage-9 input and is computed separately.
  static const double b[8] = {0.03438957868357036, 0.0, 0.0, 0.25826245556335034, 0.42093711896735372,
                              4.4053964696693102, -176.48311902429866, 172.36413340141507};
  static const double bhat[9] = {0.04909967648382, 0.0, 0.0, 0.22511122295165, 0.46946822530296,
                                 0.80657922499889, 0.0, -0.60711948917780, 0.05686113944048};
  return make_fsal_pair("vern65", 9, 6, c, a, b, 8, bhat, 9, false, false);
}

// ---- fixed-step methods and the two low-order adaptive ones (ode.nim:107-234) -------------------------
inline MethodDef make_fixed(const char* name, int stages, double order) {
  MethodDef m; m.name = name; m.stages = stages; m.order = order; m.order_int = int(order);
  for (int s = 2; s <= stages; ++s) m.a_cfac[s] = 1.0;
  return m;
}
inline MethodDef make_rk4() {  // ode.nim:180-189
  MethodDef m = make_fixed("rk4", 4, 4.0);
  m.c[2] = 0.5; m.a[2] = make_row({{1, 1.0}}); m.a_cfac[2] = 0.5;
  m.c[3] = 0.5; m.a[3] = make_row({{2, 1.0}}); m.a_cfac[3] = 0.5;
  m.c[4] = 1.0; m.a[4] = make_row({{3, 1.0}});
  m.rk4_final = true;
  return m;
}
inline MethodDef make_heun2() {  // ode.nim:107-113
  MethodDef m = make_fixed("heun2", 2, 2.0);
  m.c[2] = 1.0; m.a[2] = make_row({{1, 1.0}});
  m.b = make_row({{1, 1.0}, {2, 1.0}}); m.b_cfac = 0.5;
  return m;
}
inline MethodDef make_ralston2() {  // ode.nim:115-121
  MethodDef m = make_fixed("ralston2", 2, 2.0);
  m.c[2] = 2.0 / 3.0; m.a[2] = make_row({{1, 1.0}}); m.a_cfac[2] = 2.0 / 3.0;
  m.b = make_row({{1, 0.25}, {2, 0.75}});
  return m;
}
inline MethodDef make_kutta3() {  // ode.nim:123-130
  MethodDef m = make_fixed("kutta3", 3, 3.0);
  m.c[2] = 0.5; m.a[2] = make_row({{1, 1.0}}); m.a_cfac[2] = 0.5;
  m.c[3] = 1.0; m.a[3] = make_row({{1, -1.0}, {2, 2.0}}); m.a_chain[3] = true;  // y - dt*k1 + 2*dt*k2
  m.b = make_row({{1, 1.0 / 6.0}, {2, 2.0 / 3.0}, {3, 1.0 / 6.0}});
  return m;
}
inline MethodDef make_heun3() {  // ode.nim:132-139
  MethodDef m = make_fixed("heun3", 3, 3.0);
  m.c[2] = 1.0 / 3.0; m.a[2] = make_row({{1, 1.0}}); m.a_cfac[2] = 1.0 / 3.0;
  m.c[3] = 2.0 / 3.0; m.a[3] = make_row({{2, 1.0}}); m.a_cfac[3] = 2.0 / 3.0;
  m.b = make_row({{1, 0.25}, {3, 0.75}});
  return m;
}
inline MethodDef make_ralston3() {  // ode.nim:141-148
  MethodDef m = make_fixed("ralston3", 3, 3.0);
  m.c[2] = 1.0 / 2.0; m.a[2] = make_row({{1, 1.0}}); m.a_cfac[2] = 1.0 / 2.0;
  m.c[3] = 3.0 / 4.0; m.a[3] = make_row({{2, 1.0}}); m.a_cfac[3] = 3.0 / 4.0;
  m.b = make_row({{1, 2.0 / 9.0}, {2, 1.0 / 3.0}, {3, 4.0 / 9.0}});
  return m;
}
inline MethodDef make_ssprk3() {  // ode.nim:150-157
  MethodDef m = make_fixed("ssprk3", 3, 3.0);
  m.c[2] = 1.0; m.a[2] = make_row({{1, 1.0}});
  m.c[3] = 0.5; m.a[3] = make_row({{1, 1.0}, {2, 1.0}}); m.a_cfac[3] = 0.25;
  m.b = make_row({{1, 1.0 / 6.0}, {2, 1.0 / 6.0}, {3, 2.0 / 3.0}});
  return m;
}
inline MethodDef make_ralston4() {  // ode.nim:160-168
  MethodDef m = make_fixed("ralston4", 4, 4.0);
  m.c[2] = 0.4; m.a[2] = make_row({{1, 1.0}}); m.a_cfac[2] = 0.4;
  m.c[3] = 0.45573725; m.a[3] = make_row({{1, 0.29697761}, {2, 0.15875964}});
  m.c[4] = 1.0; m.a[4] = make_row({{1, 0.21810040}, {2, -3.05096516}, {3, 3.83286476}});
  m.b = make_row({{1, 0.17476028}, {2, -0.55148066}, {3, 1.20553560}, {4, 0.17118478}});
  return m;
}
inline MethodDef make_kutta4() {  // ode.nim:170-178
  MethodDef m = make_fixed("kutta4", 4, 4.0);
  m.c[2] = 1.0 / 3.0; m.a[2] = make_row({{1, 1.0}}); m.a_cfac[2] = 1.0 / 3.0;
  m.c[3] = 2.0 / 3.0; m.a[3] = make_row({{1, -1.0 / 3.0}, {2, 1.0}});
  m.c[4] = 1.0; m.a[4] = make_row({{1, 1.0}, {2, -1.0}, {3, 1.0}});
  m.b = make_row({{1, 1.0 / 8.0}, {2, 3.0 / 8.0}, {3, 3.0 / 8.0}, {4, 1.0 / 8.0}});
  return m;
}
inline MethodDef make_rk21() {  // ode.nim:191-210
  MethodDef m = make_fixed("rk21", 2, 2.0);
  m.adaptive = true;
  m.c[2] = 1.0; m.a[2] = make_row({{1, 1.0}});
  m.b = make_row({{1, 1.0}, {2, 1.0}}); m.b_cfac = 0.5;   // y + dt*0.5*(k1 + k2)
  m.bhat = make_row({{1, 1.0}});                           // y + dt*k1
  return m;
}
inline MethodDef make_bs32() {  // ode.nim:212-234
  MethodDef m = make_fixed("bs32", 4, 3.0);
  m.adaptive = true; m.use_fsal = true; m.fsal_out = 4; m.ynew_is_last_stage_input = true;
  m.c[2] = 0.5; m.a[2] = make_row({{1, 1.0}}); m.a_cfac[2] = 0.5;
  m.c[3] = 0.75; m.a[3] = make_row({{2, 1.0}}); m.a_cfac[3] = 0.75;
  m.c[4] = 1.0; m.a[4] = make_row({{1, 2.0 / 9.0}, {2, 1.0 / 3.0}, {3, 4.0 / 9.0}});
  m.b = m.a[4];
  m.bhat = make_row({{1, 7.0 / 24.0}, {2, 1.0 / 4.0}, {3, 1.0 / 3.0}, {4, 1.0 / 8.0}});
  return m;
}

// indexed by enum b200rk_method (include/b200rk.h)
inline const MethodDef& method_def(int id) {
  static const MethodDef tab[14] = {make_dopri54(), make_tsit54(), make_vern65(), make_rk4(), make_rk21(),
                                    make_bs32(),    make_heun2(),  make_ralston2(), make_kutta3(), make_heun3(),
                                    make_ralston3(), make_ssprk3(), make_ralston4(), make_kutta4()};
  return tab[id];
}

}  // namespace b200rk
