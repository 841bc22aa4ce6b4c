// Device-side view of a flattened mechanical system (runtime-table flavour).
//
// Two flavours of "Sys" exist with the same accessor names:
//   * RtSys  (this file)      - tables in memory, sizes known at run time: the general path.
//   * a generated `CtSys`     - `static constexpr` tables emitted by trepb_codegen.cc for one
//                               concrete system: after full unrolling every lookup folds and the
//                               frame-tree walk becomes straight-line register code.
// The math in trepb_math.cuh is written once against these accessors.
//
// Host reference for what the tables mean: trep/system.py:733-771 (frames pre-order,
// configs = dyn + kin, masses) and trep/frame.py:683-691 (cache_index == ancestor configs).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TREPB_HD __host__ __device__ __forceinline__
#define TREPB_HDN __host__ __device__
#else
#define TREPB_HD inline
#define TREPB_HDN
#endif

namespace trepb {

enum FrameKind { K_WORLD = 0, K_TX, K_TY, K_TZ, K_RX, K_RY, K_RZ, K_CONST_SE3 };
enum PotKind { P_GRAVITY = 0, P_LINEAR_SPRING, P_CONFIG_SPRING, P_NONLINEAR_CONFIG_SPRING };
enum ForceKind { F_DAMPING = 0, F_CONFIG, F_LINEAR_DAMPER, F_BODY_WRENCH, F_HYBRID_WRENCH, F_SPATIAL_WRENCH };
enum ConKind { C_DISTANCE = 0, C_POINT1D, C_PLANE };

// Status codes written per instance (never abort the batch).
enum Status { ST_OK = 0, ST_NOT_CONVERGED = -1, ST_SINGULAR = -2 };

struct RtSys {
    int nf, nd, nk, nu, nc, npot, nforce, max_depth;
    const int32_t* frame_parent;
    const int32_t* frame_kind;
    const int32_t* frame_config;
    const double* frame_value;
    const double* frame_se3;    // [nf][12]
    const double* frame_mass;   // [nf][4]
    const int32_t* cfg_frame_;  // [nq]  frame driven by config (-1: none)
    const uint8_t* dep_;        // [nf][nq]  frame f depends on config c
    const uint8_t* mass_below_; // [nf]  subtree of f (incl. f) carries mass
    const uint8_t* need_world_; // [nf]  world pose needed (point-pair consumer below)
    const uint8_t* vzero_;      // [nf]  no variable joint above f: velocity carried into f is exactly 0
    const int32_t* pot_kind_;  const int32_t* pot_i_;  const double* pot_d_;
    const int32_t* force_kind_; const int32_t* force_i_; const double* force_d_;
    const int32_t* con_kind_;  const int32_t* con_i_;  const double* con_d_;
    const int32_t* ipool_; const double* dpool_;
    double grav[3];            // sum of all Gravity potentials
    int has_gravity;
    int has_pairs;             // any point-pair element (spring/damper/constraint)
    int has_pairs_mid;         // pair elements evaluated at the midpoint (spring/damper)

    static constexpr bool kStatic = false;
    static constexpr int kUnroll = 1;   // system-sized loops stay rolled (see TREPB_UNROLL_SYS)
    static constexpr int kNPAR = 0;
    struct Params { double v[1]; };     // a generated compile-time system takes its parameters here; unused
    TREPB_HD int MAXDEPTH() const { return max_depth; }
    TREPB_HD int NF() const { return nf; }
    TREPB_HD int ND() const { return nd; }
    TREPB_HD int NK() const { return nk; }
    TREPB_HD int NQ() const { return nd + nk; }
    TREPB_HD int NU() const { return nu; }
    TREPB_HD int NC() const { return nc; }
    TREPB_HD int NPOT() const { return npot; }
    TREPB_HD int NFORCE() const { return nforce; }
    TREPB_HD int parent(int f) const { return frame_parent[f]; }
    TREPB_HD int kind(int f) const { return frame_kind[f]; }
    TREPB_HD int config(int f) const { return frame_config[f]; }
    TREPB_HD double value(int f) const { return frame_value[f]; }
    TREPB_HD double se3(int f, int k) const { return frame_se3[f * 12 + k]; }
    TREPB_HD double mass(int f, int k) const { return frame_mass[f * 4 + k]; }
    TREPB_HD int cfg_frame(int c) const { return cfg_frame_[c]; }
    TREPB_HD bool dep(int f, int c) const { return dep_[f * (nd + nk) + c] != 0; }
    TREPB_HD bool mass_below(int f) const { return mass_below_[f] != 0; }
    TREPB_HD bool need_world(int f) const { return need_world_[f] != 0; }
    TREPB_HD bool vzero(int f) const { return vzero_[f] != 0; }
    TREPB_HD bool has_mass(int f) const {
        return mass(f, 0) != 0.0 || mass(f, 1) != 0.0 || mass(f, 2) != 0.0 || mass(f, 3) != 0.0;
    }
    TREPB_HD int pot_kind(int i) const { return pot_kind_[i]; }
    TREPB_HD int pot_i(int i, int k) const { return pot_i_[i * 4 + k]; }
    TREPB_HD double pot_d(int i, int k) const { return pot_d_[i * 4 + k]; }
    TREPB_HD int force_kind(int i) const { return force_kind_[i]; }
    TREPB_HD int force_i(int i, int k) const { return force_i_[i * 4 + k]; }
    TREPB_HD double force_d(int i, int k) const { return force_d_[i * 4 + k]; }
    TREPB_HD int con_kind(int i) const { return con_kind_[i]; }
    TREPB_HD int con_i(int i, int k) const { return con_i_[i * 4 + k]; }
    TREPB_HD double con_d(int i, int k) const { return con_d_[i * 4 + k]; }
    TREPB_HD int ipool(int k) const { return ipool_[k]; }
    TREPB_HD double dpool(int k) const { return dpool_[k]; }
    TREPB_HD double gravity(int k) const { return grav[k]; }
    TREPB_HD bool gravity_on() const { return has_gravity != 0; }
    TREPB_HD bool pairs_on() const { return has_pairs != 0; }
    TREPB_HD bool pairs_mid() const { return has_pairs_mid != 0; }

    // Same tables at another address (the kernels stage the packed blob into shared memory).
    TREPB_HD RtSys rebased(const char* from, const char* to) const {
        RtSys s = *this;
#define TREPB_RB(T, p) s.p = (const T*)(to + ((const char*)p - from));
        TREPB_RB(int32_t, frame_parent) TREPB_RB(int32_t, frame_kind) TREPB_RB(int32_t, frame_config)
        TREPB_RB(double, frame_value) TREPB_RB(double, frame_se3) TREPB_RB(double, frame_mass)
        TREPB_RB(int32_t, cfg_frame_) TREPB_RB(uint8_t, dep_) TREPB_RB(uint8_t, mass_below_)
        TREPB_RB(uint8_t, need_world_) TREPB_RB(uint8_t, vzero_)
        TREPB_RB(int32_t, pot_kind_) TREPB_RB(int32_t, pot_i_) TREPB_RB(double, pot_d_)
        TREPB_RB(int32_t, force_kind_) TREPB_RB(int32_t, force_i_) TREPB_RB(double, force_d_)
        TREPB_RB(int32_t, con_kind_) TREPB_RB(int32_t, con_i_) TREPB_RB(double, con_d_)
        TREPB_RB(int32_t, ipool_) TREPB_RB(double, dpool_)
#undef TREPB_RB
        return s;
    }
};

}  // namespace trepb
