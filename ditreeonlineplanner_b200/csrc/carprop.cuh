// carprop.cuh -- the bicycle step and the per-step edge logic shared by the propagate kernels (propagate.cu) and the
// device-resident planner (planner.cu): CarEnv._update_state (car_env.py:356-396), the goal test (:341-350) and
// BasePlanner.propagate_action_sequence_env's per-step collision / goal handling (planners/base_planner.py:281-317).
#pragma once
#include "carfast.cuh"

struct Car {
  float x, y, psi, v, D, dl;
  float sn, cs;  // sin / cos of psi (shared by this step's collision test and the next step's dynamics)
};

__device__ __forceinline__ float rcp_approx(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float ex2_approx(float v) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// CarEnv._update_state (car_env.py:356-396), explicit Euler with dt = 0.02; leaves sn / cs stale
__device__ __forceinline__ void car_step(Car& c, float u0, float u1) {
  // clip to the action space (car_env.py:371; bounds car_env.py:594-597)
  u0 = fminf(fmaxf(u0, -10.0f), 10.0f);
  u1 = fminf(fmaxf(u1, -2.0f), 2.0f);
  // tanh(5 v) = 1 - 2 / (exp(10 v) + 1): saturates correctly at +-inf, absolute error ~1e-7
  const float th5v = __fmaf_rn(-2.0f, rcp_approx(ex2_approx(14.426950f * c.v) + 1.0f), 1.0f);
  const float fxd = (0.28f - 0.05f * c.v) * c.D - 0.006f * (c.v * c.v) - 0.011f * th5v;
  const float hd = 0.5f * c.dl;
  const float sh = __sinf(hd), ch = __cosf(hd);
  // cos / sin (psi + delta / 2) by angle addition
  const float cb = c.cs * ch - c.sn * sh, sb = c.sn * ch + c.cs * sh;
  const float dt = 0.02f;
  const float dpsi = c.v * 15.5f * c.dl;
  const float dv = (fxd * (1.0f / 0.043f)) * ch;
  c.x += dt * (c.v * cb);
  c.y += dt * (c.v * sb);
  c.psi += dt * dpsi;
  c.v += dt * dv;
  c.D += dt * u0;
  c.dl += dt * u1;
}

struct EdgeState {
  int first, done;
  int alive;  // 0 / 1
};

// Rare path of a step, out of line: the heading left the MUFU range, a collision decision fell inside the
// guard band, or the goal distance is within rounding of 0.5.  Re-decides everything exactly:
//   collision  : float64 code (dt_car_test, common.cuh)
//   goal       : ||p - goal|| < 0.5 as the float64 reference computes it (car_env.py:341-350); squares
//                compared, the square root taken only on the knife edge
// Returns (sn, cs, flags): sin / cos of the heading (recomputed by libm when it left the MUFU range) and
// flag bits 0 collides, 1 in goal, 2 the reference would raise IndexError.
static __device__ __noinline__ float4 edge_slow(const uint8_t* __restrict__ grid, int R, int C, float x, float y,
                                                float th, float gxf, float gyf, float sn, float cs) {
  if (!(fabsf(th) <= DT_SC_MAX)) sincosf(th, &sn, &cs);
  const int hit = dt_car_test(grid, R, C, x, y, th);
  const double ex = xsub((double)x, (double)gxf), ey = xsub((double)y, (double)gyf);
  const double d2 = xadd(xmul(ex, ex), xmul(ey, ey));
  const bool in_goal = (fabs(d2 - 0.25) < 1.0e-9) ? (__dsqrt_rn(d2) < 0.5) : (d2 < 0.25);
  return make_float4(sn, cs, __int_as_float((hit & 1) | (in_goal ? 2 : 0) | (hit & 4)), 0.f);
}

// one step of BasePlanner.propagate_action_sequence_env (planners/base_planner.py:281-317) for a live edge
template <bool kTable, bool kStop>
__device__ __forceinline__ void edge_step(Car& c, EdgeState& e, int i, float u0, float u1,
                                          const uint8_t* s_map, uint32_t s_q, const MapView& m, const QMapView& q,
                                          float gx, float gy, int* status) {
  car_step(c, u0, u1);
  dt_sincos_mufu(c.psi, c.sn, c.cs);
  // goal test (car_env.py:341-350): fp32 squared distance, trusted when clear of 0.25 by more than its error
  const float fx = c.x - gx, fy = c.y - gy;
  const float f2 = fx * fx + fy * fy;
  bool in_goal = f2 < 0.25f;
  bool rare = !(fabsf(f2 - 0.25f) > 1.0e-4f * fmaxf(1.0f, f2)) | !(fabsf(c.psi) <= DT_SC_MAX);
  // collision (planners/base_planner.py:306) on the new state
  bool coll = false;
  if (kTable) {
    bool amb;
    coll = dt_car_fast(s_q, q, c.x, c.y, c.sn, c.cs, amb);
    rare |= amb;
  } else {
    rare = true;
  }
  if (rare) {
    const float4 sl = edge_slow(s_map, m.rows, m.cols, c.x, c.y, c.psi, gx, gy, c.sn, c.cs);
    c.sn = sl.x;
    c.cs = sl.y;
    const int r = __float_as_int(sl.z);
    coll = (r & 1) != 0;
    in_goal = (r & 2) != 0;
    if (r & 4) atomicMin(status, DT_E_INDEX);
  }
  // collision ends the edge and the goal flag is then ignored (base_planner.py:306-312); goal reached: the
  // remaining actions are zeroed and the loop breaks (:314-317)
  e.first = (coll && e.first < 0) ? i : e.first;
  if (kStop) {
    e.done = (in_goal && !coll) ? i : e.done;
    e.alive = (coll || in_goal) ? 0 : 1;
  } else {
    e.done = in_goal ? i : e.done;
    e.alive = in_goal ? 0 : 1;
  }
}

