// Env-stage kernels of the hector hot path, sm_100a.
//
// Compiled with --fmad=false: the reference evaluates every expression as separate fp32 torch
// ops, so no multiply-add may be contracted here; operation order follows the reference lines
// cited at each block (paths relative to /root/reference/humanoid).
//
// Data layout: the gym tensors are read in place.  Work is tiled as ONE WARP PER 32-ENV TILE:
// every [N, row] tensor of a tile is a contiguous slab in HBM, which the warp stages into
// shared memory either with 1-D bulk async copies (TMA engine, one mbarrier) or with coalesced
// vector loads, and then each lane owns one env and reads its row from shared memory.
#include <math.h>
#include <string.h>

#include "hb_common.cuh"

namespace {

constexpr int NDOF = 10;                          // hector (hector_config.py:16)
constexpr int OBS = 5 + 3 * NDOF + 6;             // 41  (hector_env.py:219-226)
constexpr int PRIV = 5 + 3 * NDOF + 9 + 12 + 3 + 5 + 2 + 4;   // 70  (hector_env.py:195-216)
constexpr int TILE = 32;

constexpr float TWO_PI_F = 6.283185307179586f;    // float32(2*np.pi)
constexpr float PI_F = 3.141592653589793f;        // float32(np.pi)
constexpr float HALF_PI_F = 1.5707963267948966f;

struct Vec3 {
    float x, y, z;
};

__device__ __forceinline__ Vec3 cross3(const Vec3 &a, const Vec3 &b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// isaacgym.torch_utils.quat_rotate_inverse, q = (x,y,z,w): a - b + c
__device__ __forceinline__ Vec3 quat_rotate_inverse(const float q[4], const Vec3 &v) {
    const float w = q[3];
    const Vec3 u = {q[0], q[1], q[2]};
    const float s = 2.0f * (w * w) - 1.0f;
    const Vec3 cr = cross3(u, v);
    const float dot = u.x * v.x + u.y * v.y + u.z * v.z;
    Vec3 o;
    o.x = (v.x * s - (cr.x * w) * 2.0f) + (u.x * dot) * 2.0f;
    o.y = (v.y * s - (cr.y * w) * 2.0f) + (u.y * dot) * 2.0f;
    o.z = (v.z * s - (cr.z * w) * 2.0f) + (u.z * dot) * 2.0f;
    return o;
}

// isaacgym.torch_utils.quat_apply: t = 2 (u x v); v + w t + u x t
__device__ __forceinline__ Vec3 quat_apply(const float q[4], const Vec3 &v) {
    const Vec3 u = {q[0], q[1], q[2]};
    Vec3 t = cross3(u, v);
    t.x *= 2.0f, t.y *= 2.0f, t.z *= 2.0f;
    const Vec3 ut = cross3(u, t);
    return {(v.x + q[3] * t.x) + ut.x, (v.y + q[3] * t.y) + ut.y, (v.z + q[3] * t.z) + ut.z};
}

// torch.remainder(a, m) for floats
__device__ __forceinline__ float py_mod(float a, float m) {
    float r = fmodf(a, m);
    if (r != 0.0f && ((r < 0.0f) != (m < 0.0f))) r += m;
    return r;
}

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// ------------------------------------------------------------------------------------------
// a2: HectorFreeEnv.step prologue (hector_env.py:158-169, legged_robot.py:90-91)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
action_prologue_kernel(const float *__restrict__ a_in, float *__restrict__ actions, const float *__restrict__ u_delay,
                       const float *__restrict__ z_action, int total, float clip, float action_delay,
                       float action_noise) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float a = clampf(a_in[i], -clip, clip);
    const float delay = (u_delay ? u_delay[i / NDOF] : 0.0f) * action_delay;
    a = (1.0f - delay) * a + delay * actions[i];
    const float z = z_action ? z_action[i] : 0.0f;
    a = a + (action_noise * z) * a;
    actions[i] = clampf(a, -clip, clip);
}

// ------------------------------------------------------------------------------------------
// a1: LeggedRobot._compute_torques (legged_robot.py:339-355).  Two DOFs per thread so that the
// interleaved (pos,vel) pairs are one 16-byte load and everything else an 8-byte access.
// ------------------------------------------------------------------------------------------
struct PdConsts {
    float q0[HB_MAX_DOF];
    float lim[HB_MAX_DOF];
};

__global__ void __launch_bounds__(256)
pd_torque_kernel(const float4 *__restrict__ dof_state2, const float2 *__restrict__ actions2,
                 const float2 *__restrict__ kp2, const float2 *__restrict__ kd2, float2 *__restrict__ torques2,
                 int pairs, int ndof, float action_scale, const __grid_constant__ PdConsts c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pairs) return;
    const float4 s = dof_state2[i];          // q0 qd0 q1 qd1
    const float2 a = actions2[i], kp = kp2[i], kd = kd2[i];
    const int j = (2 * i) % ndof;            // ndof even: a pair never straddles two envs
    float2 t;
    t.x = kp.x * ((a.x * action_scale + c.q0[j]) - s.x) - kd.x * s.y;
    t.y = kp.y * ((a.y * action_scale + c.q0[j + 1]) - s.z) - kd.y * s.w;
    t.x = clampf(t.x, -c.lim[j], c.lim[j]);
    t.y = clampf(t.y, -c.lim[j + 1], c.lim[j + 1]);
    torques2[i] = t;
}

__global__ void __launch_bounds__(256)
pd_torque_scalar_kernel(const float *__restrict__ dof_state, const float *__restrict__ actions,
                        const float *__restrict__ kp, const float *__restrict__ kd, float *__restrict__ torques,
                        int total, int ndof, float action_scale, const __grid_constant__ PdConsts c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int j = i % ndof;
    const float t = kp[i] * ((actions[i] * action_scale + c.q0[j]) - dof_state[2 * i]) - kd[i] * dof_state[2 * i + 1];
    torques[i] = clampf(t, -c.lim[j], c.lim[j]);
}

// ------------------------------------------------------------------------------------------
// a3-a9: post_physics_step.  One CTA = one 32-env tile, four specialised warps, lane = env.
//
//   warp 0 "base"    root quaternion work: base_lin_vel / base_ang_vel / projected_gravity / euler,
//                    command resampling + heading command, push; rewards base_acc, orientation,
//                    tracking_ang_vel, tracking_lin_vel
//   warp 1 "joints"  dof / action terms: action_smoothness, default_joint_pos, dof_acc, dof_vel, torques
//   warp 2 "feet"    gait phase, contacts, termination + time-out -> reset flag; base_height, collision,
//                    feet_air_time, feet_clearance, feet_contact_forces, feet_contact_number,
//                    feet_distance, foot_slip, knee_distance
//   warp 3 "ledger"  episode sums; after the barrier: alphabetical reward accumulation
//                    (legged_robot.py:216-234), reset compaction and episode means
//
// Every warp stays convergent (all lanes run the same code on different envs), the serial chain
// per env is a quarter of the fused version, and each warp only fetches its own slice of code.
// Phase A (rewards, reset flag) | barrier | phase B (reset_idx effects, newest frames, write-back)
// | barrier | coalesced store of the frames.
// ------------------------------------------------------------------------------------------
struct TileLayout {          // float offsets into the CTA's shared-memory tile
    int root, dof, contact, actions, last_actions, last_last_actions, last_dof_vel, torques, last_root_vel,
        commands, z_obs, terms, gait_s, gait_c, flags, so, sp, total;
};
constexpr int PRIV_PAD = PRIV + 1;     // odd row stride: conflict-free lane-per-row writes

__host__ __device__ inline TileLayout make_layout(int nbody, bool with_noise) {
    TileLayout L;
    int o = 0;
    L.root = o, o += TILE * 13;
    L.dof = o, o += TILE * NDOF * 2;
    L.contact = o, o += TILE * nbody * 3;
    L.actions = o, o += TILE * NDOF;
    L.last_actions = o, o += TILE * NDOF;
    L.last_last_actions = o, o += TILE * NDOF;
    L.last_dof_vel = o, o += TILE * NDOF;
    L.torques = o, o += TILE * NDOF;
    L.last_root_vel = o, o += TILE * 6;
    L.commands = o, o += TILE * 4;
    L.z_obs = o, o += with_noise ? TILE * OBS : 0;
    L.terms = o, o += HB_NUM_REWARDS * TILE;
    L.gait_s = o, o += TILE;
    L.gait_c = o, o += TILE;
    L.flags = o, o += TILE;
    L.so = o, o += TILE * OBS;
    L.sp = o, o += TILE * PRIV_PAD;
    L.total = o;
    return L;
}

// all four warps of the tile (named barrier 1: the roles meet at different program points)
__device__ __forceinline__ void tile_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ float norm3(const float *f) { return sqrtf((f[0] * f[0] + f[1] * f[1]) + f[2] * f[2]); }

// torch.remainder(a, 2*pi) for |a| < 2*pi (every angle here), general fmodf path kept out of line
__device__ __noinline__ float py_mod_slow(float a, float m) { return py_mod(a, m); }
__device__ __forceinline__ float mod_two_pi(float a) {
    if (fabsf(a) < TWO_PI_F) return (a < 0.0f) ? a + TWO_PI_F : a;
    return py_mod_slow(a, TWO_PI_F);
}
// get_euler_xyz_tensor (envs/base/legged_robot.py:50-55) over isaacgym get_euler_xyz
__device__ __noinline__ Vec3 euler_xyz_wrapped_nl(float x, float y, float z, float w) {
    float roll = atan2f(2.0f * (w * x + y * z), ((w * w - x * x) - y * y) + z * z);
    const float sinp = 2.0f * (w * y - z * x);
    float pitch;
    if (fabsf(sinp) >= 1.0f) {
        const float sg = (sinp > 0.0f) ? 1.0f : ((sinp < 0.0f) ? -1.0f : 0.0f);
        pitch = HALF_PI_F * sg;
    } else {
        pitch = asinf(sinp);
    }
    float yaw = atan2f(2.0f * (w * z + x * y), ((w * w + x * x) - y * y) - z * z);
    roll = mod_two_pi(roll), pitch = mod_two_pi(pitch), yaw = mod_two_pi(yaw);
    if (roll > PI_F) roll -= TWO_PI_F;
    if (pitch > PI_F) pitch -= TWO_PI_F;
    if (yaw > PI_F) yaw -= TWO_PI_F;
    return {roll, pitch, yaw};
}

template <bool kBulk>
__global__ void __launch_bounds__(4 * TILE)
post_physics_kernel(const __grid_constant__ hb_env_params p, const __grid_constant__ hb_env_buffers b,
                    const __grid_constant__ hb_env_noise nz, float *__restrict__ obs_new,
                    float *__restrict__ priv_new, int stages, int32_t *host_count) {
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int s_is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.num_envs;
    const int env0 = blockIdx.x * TILE;
    const int nv = min(TILE, N - env0);
    const int env = env0 + lane;
    const bool valid = lane < nv;
    const int ln = valid ? lane : 0;
    const bool do_step = stages & HB_STAGE_STEP;
    const bool derive = do_step || (stages & HB_STAGE_DERIVE);
    const bool emit_obs = do_step || (stages & HB_STAGE_OBS);
    const bool with_noise = p.add_noise && nz.z_obs != nullptr;
    const TileLayout L = make_layout(p.num_bodies, with_noise);
    const int crow = p.num_bodies * 3;
    const float dt = p.dt;
    const float clip = p.clip_observations;
    float *terms = sm + L.terms;
    float *so = sm + L.so, *sp = sm + L.sp;
    int *flags = reinterpret_cast<int *>(sm + L.flags);

    // ---------------- stage the tile's input slabs into shared memory ----------------
    const size_t e0 = env0;
    const bool full = (nv == TILE);
    const bool bulk = kBulk && full;
    if (bulk && threadIdx.x == 0) {
        hb::mbar_init(&bar, 1);
        hb::fence_mbar_init();
        hb::mbar_expect_tx(&bar, TILE * 4u * (13 + 2 * NDOF + crow + 5 * NDOF + 6 + 4 + (with_noise ? OBS : 0)));
    }
    auto stage = [&](const float *src, int dst, int row) {
        if (bulk) {
            if (threadIdx.x == 0) hb::bulk_g2s(sm + dst, src, TILE * 4u * row, &bar);
        } else {
            const int count = nv * row;
            for (int k = threadIdx.x; k < count; k += blockDim.x) sm[dst + k] = __ldg(src + k);
        }
    };
    stage(b.root_states + e0 * 13, L.root, 13);
    stage(b.dof_state + e0 * NDOF * 2, L.dof, NDOF * 2);
    stage(b.contact_forces + e0 * crow, L.contact, crow);
    stage(b.actions + e0 * NDOF, L.actions, NDOF);
    stage(b.last_actions + e0 * NDOF, L.last_actions, NDOF);
    stage(b.last_last_actions + e0 * NDOF, L.last_last_actions, NDOF);
    stage(b.last_dof_vel + e0 * NDOF, L.last_dof_vel, NDOF);
    stage(b.torques + e0 * NDOF, L.torques, NDOF);
    stage(b.last_root_vel + e0 * 6, L.last_root_vel, 6);
    stage(b.commands + e0 * 4, L.commands, 4);
    if (with_noise) stage(nz.z_obs + e0 * OBS, L.z_obs, OBS);

    // ---------------- per-warp global loads that overlap the staging ----------------
    long long ep_len = 0;                       // warps 0 and 2
    float sums[HB_NUM_REWARDS];                 // warp 3
    float foot_pos[2][3], foot_vel[2][3], knee_xy[2][2], air[2], fh[2], lz[2];   // warp 2
    bool last_ct[2];
    float push_f[2], push_t[3];                 // warp 0
    if (valid) {
        if (warp == 0 || warp == 2) ep_len = b.episode_length_buf[env];
        if (warp == 0) {
            push_f[0] = b.rand_push_force[env * 3], push_f[1] = b.rand_push_force[env * 3 + 1];
#pragma unroll
            for (int k = 0; k < 3; ++k) push_t[k] = b.rand_push_torque[env * 3 + k];
        } else if (warp == 2) {
            const float *rs = b.rigid_state + (size_t)env * p.num_bodies * 13;
#pragma unroll
            for (int f = 0; f < 2; ++f) {
                const float *r = rs + p.feet[f] * 13;
#pragma unroll
                for (int k = 0; k < 3; ++k) foot_pos[f][k] = __ldg(r + k), foot_vel[f][k] = __ldg(r + 7 + k);
                const float *kr = rs + p.knees[f] * 13;
                knee_xy[f][0] = __ldg(kr), knee_xy[f][1] = __ldg(kr + 1);
                air[f] = b.feet_air_time[env * 2 + f];
                fh[f] = b.feet_height[env * 2 + f];
                lz[f] = b.last_feet_z[env * 2 + f];
                last_ct[f] = b.last_contacts[env * 2 + f] != 0;
            }
        } else if (warp == 3) {
#pragma unroll
            for (int k = 0; k < HB_NUM_REWARDS; ++k) sums[k] = b.episode_sums[(size_t)k * N + env];
        }
    }
    __syncthreads();                             // mbarrier init / plain staging visible to every warp
    if (bulk) hb::mbar_wait(&bar, 0);

    // =========================================================================================
    if (warp == 0) {
        float root[13], cmd[4];
#pragma unroll
        for (int k = 0; k < 13; ++k) root[k] = sm[L.root + ln * 13 + k];
#pragma unroll
        for (int k = 0; k < 4; ++k) cmd[k] = sm[L.commands + ln * 4 + k];
        Vec3 lin = {0.f, 0.f, 0.f}, ang = {0.f, 0.f, 0.f}, grav = {0.f, 0.f, 0.f}, eul = {0.f, 0.f, 0.f};
        const float *quat = root + 3;
        if (derive) {        // legged_robot.py:127-135 (and _init_buffers :452,477-479)
            lin = quat_rotate_inverse(quat, {root[7], root[8], root[9]});
            ang = quat_rotate_inverse(quat, {root[10], root[11], root[12]});
            grav = quat_rotate_inverse(quat, {0.0f, 0.0f, -1.0f});
        } else if (valid) {  // reset / observation-only pass: keep the stored values
            lin = {b.base_lin_vel[env * 3], b.base_lin_vel[env * 3 + 1], b.base_lin_vel[env * 3 + 2]};
            ang = {b.base_ang_vel[env * 3], b.base_ang_vel[env * 3 + 1], b.base_ang_vel[env * 3 + 2]};
            grav = {b.projected_gravity[env * 3], b.projected_gravity[env * 3 + 1], b.projected_gravity[env * 3 + 2]};
        }
        eul = euler_xyz_wrapped_nl(quat[0], quat[1], quat[2], quat[3]);
        if (do_step) {
            ep_len += 1;
            // -------- _post_physics_step_callback, legged_robot.py:303-335 --------
            if (valid && (ep_len % p.resample_interval) == 0 && nz.u_cmd) {
                const float *u = nz.u_cmd + (size_t)env * 3;
                cmd[0] = p.cmd_span[0] * u[0] + p.cmd_lo[0];
                cmd[1] = p.cmd_span[1] * u[1] + p.cmd_lo[1];
                cmd[3] = p.cmd_span[2] * u[2] + p.cmd_lo[2];
                const float keep = (sqrtf(cmd[0] * cmd[0] + cmd[1] * cmd[1]) > 0.2f) ? 1.0f : 0.0f;
                cmd[0] *= keep, cmd[1] *= keep;
            }
            if (p.heading_command) {
                const Vec3 fwd = quat_apply(quat, {1.0f, 0.0f, 0.0f});
                const float heading = atan2f(fwd.y, fwd.x);
                float e = mod_two_pi(cmd[3] - heading);                 // wrap_to_pi, utils/math.py:46-49
                e = e - TWO_PI_F * ((e > PI_F) ? 1.0f : 0.0f);
                cmd[2] = clampf(0.5f * e, -1.0f, 1.0f);
            }
            if ((stages & HB_STAGE_PUSH) && valid && nz.u_push) {       // _push_robots, hector_env.py:53-68
                const float *u = nz.u_push + (size_t)env * 5;
                push_f[0] = p.push_lin_span * u[0] + p.push_lin_lo;
                push_f[1] = p.push_lin_span * u[1] + p.push_lin_lo;
                root[7] = push_f[0], root[8] = push_f[1];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    push_t[k] = p.push_ang_span * u[2 + k] + p.push_ang_lo;
                    root[10 + k] = push_t[k];
                }
                b.rand_push_force[env * 3] = push_f[0], b.rand_push_force[env * 3 + 1] = push_f[1];
#pragma unroll
                for (int k = 0; k < 3; ++k) b.rand_push_torque[env * 3 + k] = push_t[k];
            }
            {   // base_acc, hector_env.py:385-392 (root velocity after the push)
                float ss = 0.f;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const float d = sm[L.last_root_vel + ln * 6 + k] - root[7 + k];
                    ss += d * d;
                }
                terms[HB_R_BASE_ACC * TILE + lane] = expf(-sqrtf(ss) * 3.0f);
            }
            {   // orientation, :341-348
                const float a = expf(-(fabsf(eul.x) + fabsf(eul.y)) * 10.0f);
                const float g = expf(-sqrtf(grav.x * grav.x + grav.y * grav.y) * 20.0f);
                terms[HB_R_ORIENTATION * TILE + lane] = (a + g) / 2.0f;
            }
            {   // tracking_ang_vel :435-443, tracking_lin_vel :426-433
                const float ea = (cmd[2] - ang.z) * (cmd[2] - ang.z);
                terms[HB_R_TRACKING_ANG_VEL * TILE + lane] = expf(-ea * p.tracking_sigma);
                const float ex = cmd[0] - lin.x, ey = cmd[1] - lin.y;
                terms[HB_R_TRACKING_LIN_VEL * TILE + lane] = expf(-(ex * ex + ey * ey) * p.tracking_sigma);
            }
        }
        tile_barrier();                                           // ---- barrier 1 ----
        const bool reset = valid && (flags[lane] & 1);
        const float gs = reset ? 0.0f : sm[L.gait_s + lane], gc = reset ? 1.0f : sm[L.gait_c + lane];
        if (reset) {         // _reset_root_states + _resample_commands + the gravity/euler fix-up (:373-396,321-335,211-214)
            const float *u = nz.u_reset + (size_t)env * 15;
#pragma unroll
            for (int k = 0; k < 13; ++k) root[k] = p.base_init_state[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) root[k] += b.env_origins[env * 3 + k];
            if (p.custom_origins) {
                root[0] += p.reset_xy_span * u[10] + p.reset_xy_lo;
                root[1] += p.reset_xy_span * u[11] + p.reset_xy_lo;
            }
            cmd[0] = p.cmd_span[0] * u[12] + p.cmd_lo[0];
            cmd[1] = p.cmd_span[1] * u[13] + p.cmd_lo[1];
            cmd[3] = p.cmd_span[2] * u[14] + p.cmd_lo[2];
            const float keep = (sqrtf(cmd[0] * cmd[0] + cmd[1] * cmd[1]) > 0.2f) ? 1.0f : 0.0f;
            cmd[0] *= keep, cmd[1] *= keep;
            grav = {p.reset_gravity[0], p.reset_gravity[1], p.reset_gravity[2]};
            eul = {p.reset_euler[0], p.reset_euler[1], p.reset_euler[2]};
        }
        if (valid) {
            if (reset) {
#pragma unroll
                for (int k = 0; k < 13; ++k) b.root_states[(size_t)env * 13 + k] = root[k];
            } else if (stages & HB_STAGE_PUSH) {
                b.root_states[(size_t)env * 13 + 7] = root[7], b.root_states[(size_t)env * 13 + 8] = root[8];
#pragma unroll
                for (int k = 0; k < 3; ++k) b.root_states[(size_t)env * 13 + 10 + k] = root[10 + k];
            }
            if (derive) {
                b.base_lin_vel[env * 3] = lin.x, b.base_lin_vel[env * 3 + 1] = lin.y, b.base_lin_vel[env * 3 + 2] = lin.z;
                b.base_ang_vel[env * 3] = ang.x, b.base_ang_vel[env * 3 + 1] = ang.y, b.base_ang_vel[env * 3 + 2] = ang.z;
            }
            if (do_step) {
#pragma unroll
                for (int k = 0; k < 6; ++k) b.last_root_vel[(size_t)env * 6 + k] = root[7 + k];
            }
            b.projected_gravity[env * 3] = grav.x, b.projected_gravity[env * 3 + 1] = grav.y, b.projected_gravity[env * 3 + 2] = grav.z;
            b.base_euler_xyz[env * 3] = eul.x, b.base_euler_xyz[env * 3 + 1] = eul.y, b.base_euler_xyz[env * 3 + 2] = eul.z;
#pragma unroll
            for (int k = 0; k < 4; ++k) b.commands[(size_t)env * 4 + k] = cmd[k];
        }
        if (emit_obs) {      // command input, base velocities, euler, root position, push (hector_env.py:186-226)
            constexpr int B0 = 5 + 3 * NDOF;
            float o[11];
            o[0] = gs, o[1] = gc;
            o[2] = cmd[0] * p.obs_lin_vel, o[3] = cmd[1] * p.obs_lin_vel, o[4] = cmd[2] * p.obs_ang_vel;
            o[5] = ang.x * p.obs_ang_vel, o[6] = ang.y * p.obs_ang_vel, o[7] = ang.z * p.obs_ang_vel;
            o[8] = eul.x * p.obs_quat, o[9] = eul.y * p.obs_quat, o[10] = eul.z * p.obs_quat;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                so[lane * OBS + k] = clampf(o[k], -clip, clip);        // noise scale of the command slots is 0
                sp[lane * PRIV_PAD + k] = clampf(o[k], -clip, clip);
            }
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                float v = o[5 + k];
                sp[lane * PRIV_PAD + B0 + 3 + k] = clampf(v, -clip, clip);
                if (with_noise) v = v + (sm[L.z_obs + ln * OBS + B0 + k] * p.noise_scale_vec[B0 + k]) * p.noise_level;
                so[lane * OBS + B0 + k] = clampf(v, -clip, clip);
            }
            sp[lane * PRIV_PAD + B0] = clampf(lin.x * p.obs_lin_vel, -clip, clip);
            sp[lane * PRIV_PAD + B0 + 1] = clampf(lin.y * p.obs_lin_vel, -clip, clip);
            sp[lane * PRIV_PAD + B0 + 2] = clampf(lin.z * p.obs_lin_vel, -clip, clip);
#pragma unroll
            for (int k = 0; k < 3; ++k) sp[lane * PRIV_PAD + B0 + 21 + k] = clampf(root[k], -clip, clip);
            sp[lane * PRIV_PAD + B0 + 24] = clampf(push_f[0], -clip, clip);
            sp[lane * PRIV_PAD + B0 + 25] = clampf(push_f[1], -clip, clip);
#pragma unroll
            for (int k = 0; k < 3; ++k) sp[lane * PRIV_PAD + B0 + 26 + k] = clampf(push_t[k], -clip, clip);
        }
    } else if (warp == 1) {
        // ================================ joints ================================
        float q[NDOF], qd[NDOF], act[NDOF], lact[NDOF], llact[NDOF], ldv[NDOF];
#pragma unroll
        for (int j = 0; j < NDOF; ++j) {
            q[j] = sm[L.dof + ln * NDOF * 2 + 2 * j];
            qd[j] = sm[L.dof + ln * NDOF * 2 + 2 * j + 1];
            act[j] = sm[L.actions + ln * NDOF + j];
            lact[j] = sm[L.last_actions + ln * NDOF + j];
            llact[j] = sm[L.last_last_actions + ln * NDOF + j];
            ldv[j] = sm[L.last_dof_vel + ln * NDOF + j];
        }
        if (do_step) {
            float t1 = 0.f, t2 = 0.f, t3 = 0.f, ss = 0.f, acc = 0.f, vel = 0.f, tq = 0.f, d[NDOF];
#pragma unroll
            for (int j = 0; j < NDOF; ++j) {
                const float d1 = lact[j] - act[j];                     // action_smoothness, hector_env.py:529-539
                t1 += d1 * d1;
                const float d2 = (act[j] + llact[j]) - 2.0f * lact[j];
                t2 += d2 * d2;
                t3 += fabsf(act[j]);
                d[j] = q[j] - p.default_dof_pos[j];                    // default_joint_pos, :357-367
                ss += d[j] * d[j];
                const float a = (ldv[j] - qd[j]) / dt;                 // dof_acc :515-520
                acc += a * a;
                vel += qd[j] * qd[j];                                  // dof_vel :508-513
                const float tau = sm[L.torques + ln * NDOF + j];       // torques :501-506
                tq += tau * tau;
            }
            terms[HB_R_ACTION_SMOOTHNESS * TILE + lane] = (t1 + t2) + 0.05f * t3;
            float yr = sqrtf(d[0] * d[0] + d[1] * d[1]) + sqrtf(d[5] * d[5] + d[6] * d[6]);
            yr = clampf(yr - 0.1f, 0.0f, 50.0f);
            terms[HB_R_DEFAULT_JOINT_POS * TILE + lane] = expf(-yr * 100.0f) - 0.01f * sqrtf(ss);
            terms[HB_R_DOF_ACC * TILE + lane] = acc;
            terms[HB_R_DOF_VEL * TILE + lane] = vel;
            terms[HB_R_TORQUES * TILE + lane] = tq;
        }
        tile_barrier();                                           // ---- barrier 1 ----
        const bool reset = valid && (flags[lane] & 1);
        if (reset) {         // _reset_dofs (legged_robot.py:358-372) + buffer zeroing (:186-191)
            const float *u = nz.u_reset + (size_t)env * 15;
#pragma unroll
            for (int j = 0; j < NDOF; ++j) {
                q[j] = p.default_dof_pos[j] + (p.reset_dof_span * u[j] + p.reset_dof_lo);
                qd[j] = 0.0f;
                b.dof_state[((size_t)env * NDOF + j) * 2] = q[j];
                b.dof_state[((size_t)env * NDOF + j) * 2 + 1] = 0.0f;
                act[j] = lact[j] = llact[j] = 0.0f;
                b.actions[(size_t)env * NDOF + j] = 0.0f;
            }
        }
        if (valid) {         // legged_robot.py:146-148
#pragma unroll
            for (int j = 0; j < NDOF; ++j) {
                b.last_last_actions[(size_t)env * NDOF + j] = do_step ? lact[j] : llact[j];
                b.last_actions[(size_t)env * NDOF + j] = do_step ? act[j] : lact[j];
                b.last_dof_vel[(size_t)env * NDOF + j] = (do_step || reset) ? qd[j] : ldv[j];
            }
        }
        if (emit_obs) {
#pragma unroll
            for (int j = 0; j < NDOF; ++j) {
                float v0 = (q[j] - p.default_dof_pos[j]) * p.obs_dof_pos, v1 = qd[j] * p.obs_dof_vel;
                sp[lane * PRIV_PAD + 5 + j] = clampf(v0, -clip, clip);
                sp[lane * PRIV_PAD + 5 + NDOF + j] = clampf(v1, -clip, clip);
                sp[lane * PRIV_PAD + 5 + 2 * NDOF + j] = clampf(act[j], -clip, clip);
                if (with_noise) {
                    v0 = v0 + (sm[L.z_obs + ln * OBS + 5 + j] * p.noise_scale_vec[5 + j]) * p.noise_level;
                    v1 = v1 + (sm[L.z_obs + ln * OBS + 5 + NDOF + j] * p.noise_scale_vec[5 + NDOF + j]) * p.noise_level;
                }
                float v2 = act[j];
                if (with_noise)
                    v2 = v2 + (sm[L.z_obs + ln * OBS + 5 + 2 * NDOF + j] * p.noise_scale_vec[5 + 2 * NDOF + j]) * p.noise_level;
                so[lane * OBS + 5 + j] = clampf(v0, -clip, clip);
                so[lane * OBS + 5 + NDOF + j] = clampf(v1, -clip, clip);
                so[lane * OBS + 5 + 2 * NDOF + j] = clampf(v2, -clip, clip);
            }
        }
    } else if (warp == 2) {
        // ================================ feet / contacts / gait ================================
        const float *cf = sm + L.contact + ln * crow;
        float foot_f[2][3];
#pragma unroll
        for (int f = 0; f < 2; ++f)
#pragma unroll
            for (int k = 0; k < 3; ++k) foot_f[f][k] = cf[p.feet[f] * 3 + k];
        const float root_z = sm[L.root + ln * 13 + 2];
        bool reset = false, time_out = false;
        float gs = 0.0f, gc = 1.0f, st[2] = {1.0f, 1.0f};
        const bool ct[2] = {foot_f[0][2] > 5.0f, foot_f[1][2] > 5.0f};
        if (do_step) {
            ep_len += 1;
            // -------- check_termination, legged_robot.py:155-160 --------
            float coll = 0.f;
#pragma unroll 1
            for (int i = 0; i < p.n_term; ++i) reset |= norm3(cf + p.term_bodies[i] * 3) > 1.0f;
#pragma unroll 1
            for (int i = 0; i < p.n_pen; ++i) coll += (norm3(cf + p.pen_bodies[i] * 3) > 0.1f) ? 1.0f : 0.0f;
            time_out = ep_len > (long long)p.max_episode_length;
            reset |= time_out;
            terms[HB_R_COLLISION * TILE + lane] = coll;                 // :522-527
        }
        {   // -------- gait phase, hector_env.py:70-88 --------
            const float arg = TWO_PI_F * (((float)ep_len * dt) / p.cycle_time);
            sincosf(arg, &gs, &gc);
            st[0] = (gs >= 0.0f) ? 1.0f : 0.0f, st[1] = (gs < 0.0f) ? 1.0f : 0.0f;
            if (fabsf(gs) < 0.1f) st[0] = st[1] = 1.0f;
        }
        if (do_step) {
            {   // base_height, :369-383
                const float ground = (foot_pos[0][2] * st[0] + foot_pos[1][2] * st[1]) / (st[0] + st[1]);
                const float h = root_z - (ground - 0.05f);
                terms[HB_R_BASE_HEIGHT * TILE + lane] = expf(-fabsf(h - p.base_height_target) * 100.0f);
            }
            float r_air = 0.f, r_clr = 0.f, cfz = 0.f, num = 0.f, slip = 0.f;
#pragma unroll
            for (int f = 0; f < 2; ++f) {
                {   // feet_air_time, :315-329 (stateful)
                    const bool filt = ct[f] || (st[f] != 0.0f) || last_ct[f];
                    last_ct[f] = ct[f];
                    const bool first = (air[f] > 0.0f) && filt;
                    air[f] += dt;
                    r_air += clampf(air[f], 0.0f, 0.5f) * (first ? 1.0f : 0.0f);
                    air[f] *= filt ? 0.0f : 1.0f;
                }
                {   // feet_clearance, :445-466 (stateful)
                    const float z = foot_pos[f][2] - 0.05f;
                    fh[f] += z - lz[f];
                    lz[f] = z;
                    const float hit = (fabsf(fh[f] - p.target_feet_height) < 0.01f) ? 1.0f : 0.0f;
                    r_clr += hit * (1.0f - st[f]);
                    fh[f] *= ct[f] ? 0.0f : 1.0f;
                }
                // feet_contact_forces :350-355, feet_contact_number :331-339, foot_slip :303-313
                cfz += clampf(norm3(foot_f[f]) - p.max_contact_force, 0.0f, 400.0f);
                num += (ct[f] == (st[f] != 0.0f)) ? 1.0f : -0.3f;
                const float spd = sqrtf(sqrtf(foot_vel[f][0] * foot_vel[f][0] + foot_vel[f][1] * foot_vel[f][1]));
                slip += spd * (ct[f] ? 1.0f : 0.0f);
            }
            terms[HB_R_FEET_AIR_TIME * TILE + lane] = r_air;
            terms[HB_R_FEET_CLEARANCE * TILE + lane] = r_clr;
            terms[HB_R_FEET_CONTACT_FORCES * TILE + lane] = cfz;
            terms[HB_R_FEET_CONTACT_NUMBER * TILE + lane] = num / 2.0f;
            terms[HB_R_FOOT_SLIP * TILE + lane] = slip;
            {   // feet_distance :277-287, knee_distance :290-300
                float dx = foot_pos[0][0] - foot_pos[1][0], dy = foot_pos[0][1] - foot_pos[1][1];
                float d = sqrtf(dx * dx + dy * dy);
                float dmin = clampf(d - p.min_dist, -0.5f, 0.0f), dmax = clampf(d - p.max_dist, 0.0f, 0.5f);
                terms[HB_R_FEET_DISTANCE * TILE + lane] = (expf(-fabsf(dmin) * 100.0f) + expf(-fabsf(dmax) * 100.0f)) / 2.0f;
                dx = knee_xy[0][0] - knee_xy[1][0], dy = knee_xy[0][1] - knee_xy[1][1];
                d = sqrtf(dx * dx + dy * dy);
                dmin = clampf(d - p.min_dist, -0.5f, 0.0f), dmax = clampf(d - p.max_dist / 2.0f, 0.0f, 0.5f);
                terms[HB_R_KNEE_DISTANCE * TILE + lane] = (expf(-fabsf(dmin) * 100.0f) + expf(-fabsf(dmax) * 100.0f)) / 2.0f;
            }
        }
        if (stages & HB_STAGE_RESET_ALL) reset = true;
        if ((stages & HB_STAGE_RESET_MASK) && valid) reset = reset || (b.reset_buf[env] != 0);   // reset_idx(env_ids)
        reset = reset && valid;
        flags[lane] = (reset ? 1 : 0) | (time_out ? 2 : 0);
        sm[L.gait_s + lane] = gs, sm[L.gait_c + lane] = gc;
        tile_barrier();                                           // ---- barrier 1 ----
        if (reset) {
            ep_len = 0, air[0] = air[1] = 0.0f;
            st[0] = st[1] = 1.0f;                                   // phase 0: sin = 0 -> double support
        }
        if (valid) {
            b.feet_air_time[env * 2] = air[0], b.feet_air_time[env * 2 + 1] = air[1];
            b.episode_length_buf[env] = ep_len;
            b.reset_buf[env] = reset ? 1 : 0;
            if (do_step) {
                b.time_out_buf[env] = time_out ? 1 : 0;
#pragma unroll
                for (int f = 0; f < 2; ++f) {
                    b.last_contacts[env * 2 + f] = last_ct[f] ? 1 : 0;
                    b.feet_height[env * 2 + f] = fh[f];
                    b.last_feet_z[env * 2 + f] = lz[f];
                }
            }
        }
        if (emit_obs) {
            constexpr int B0 = 5 + 3 * NDOF;
#pragma unroll
            for (int f = 0; f < 2; ++f)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    sp[lane * PRIV_PAD + B0 + 9 + f * 3 + k] = clampf(foot_pos[f][k], -clip, clip);
                    sp[lane * PRIV_PAD + B0 + 15 + f * 3 + k] = clampf(foot_vel[f][k], -clip, clip);
                }
            sp[lane * PRIV_PAD + B0 + 29] = clampf(valid ? b.env_frictions[env] : 0.0f, -clip, clip);
            sp[lane * PRIV_PAD + B0 + 30] = clampf((valid ? b.body_mass[env] : 0.0f) / 30.0f, -clip, clip);
            sp[lane * PRIV_PAD + B0 + 31] = st[0], sp[lane * PRIV_PAD + B0 + 32] = st[1];
            sp[lane * PRIV_PAD + B0 + 33] = ct[0] ? 1.0f : 0.0f, sp[lane * PRIV_PAD + B0 + 34] = ct[1] ? 1.0f : 0.0f;
        }
    } else {
        // ================================ ledger ================================
        tile_barrier();                                           // ---- barrier 1 ----
        const bool reset = valid && (flags[lane] & 1);
        if (do_step) {       // compute_reward, legged_robot.py:216-234: alphabetical accumulation
            float rew = 0.0f;
#pragma unroll
            for (int k = 0; k < HB_NUM_REWARDS; ++k) {
                if (p.reward_scale[k] != 0.0f) {
                    const float r = terms[k * TILE + lane] * p.reward_scale[k];
                    rew += r;
                    sums[k] += r;
                }
            }
            if (p.only_positive_rewards) rew = fmaxf(rew, 0.0f);
            if (valid) b.rew_buf[env] = rew;
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, reset);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < HB_NUM_REWARDS; ++k) {
            terms[k * TILE + lane] = reset ? sums[k] : 0.0f;        // episode sums of the envs being reset (:198-201)
            if (reset) sums[k] = 0.0f;
            if (valid) b.episode_sums[(size_t)k * N + env] = sums[k];
        }
        __syncwarp();
        // ---------------- ordered compaction of the reset ids + episode means ----------------
        // Each tile publishes its ballot and the per-term sums of its reset envs; the last CTA to
        // finish scans the ballots in tile order (ascending env ids, like reset_buf.nonzero()).
        if (ballot && lane < HB_NUM_REWARDS) {
            float v = 0.0f;
#pragma unroll 1
            for (int e = 0; e < TILE; ++e) v += terms[lane * TILE + ((e + lane) & 31)];
            b.scratch_partials[(size_t)blockIdx.x * HB_NUM_REWARDS + lane] = v;
        }
        if (lane == 0) b.scratch_ballots[blockIdx.x] = ballot;
        __threadfence();                          // publish ballot + partials before this CTA takes its ticket
    }

    // ---------------- coalesced store of the newest frames into the last slot of the stacked buffers ----------------
    __syncthreads();                                              // ---- barrier 2 ----
    if (threadIdx.x == 0) s_is_last = (atomicAdd(b.scratch_ticket, 1u) == gridDim.x - 1);
    if (emit_obs) {
        const int ostride = p.frame_stack * OBS, pstride = p.c_frame_stack * PRIV;
        for (int i = threadIdx.x; i < nv * OBS; i += blockDim.x) {
            const int r = i / OBS, c = i - r * OBS;
            obs_new[(size_t)(env0 + r) * ostride + (ostride - OBS) + c] = so[i];
        }
        for (int i = threadIdx.x; i < nv * PRIV; i += blockDim.x) {
            const int r = i / PRIV, c = i - r * PRIV;
            priv_new[(size_t)(env0 + r) * pstride + (pstride - PRIV) + c] = sp[r * PRIV_PAD + c];
        }
    }
    __syncthreads();
    if (!s_is_last) return;

    // ---------------- last CTA: ordered compaction of the reset ids + episode means ----------------
    // Every tile has published its ballot and the per-term sums of its reset envs.  The whole CTA
    // scans the ballots in tile order (ascending env ids, like reset_buf.nonzero()): thread t owns a
    // contiguous run of tiles, so ids and the fp64 partial sums are combined in a fixed order.
    __threadfence();
    int *scan = reinterpret_cast<int *>(sm);                     // [128] tile-run counts (tile memory is free now)
    double *red = reinterpret_cast<double *>(sm + 256);          // [HB_NUM_REWARDS][128]
    const int tiles = gridDim.x;
    const int per = (tiles + blockDim.x - 1) / blockDim.x;
    const int t_lo = min((int)threadIdx.x * per, tiles), t_hi = min(t_lo + per, tiles);
    int cnt = 0;
    for (int t = t_lo; t < t_hi; ++t) cnt += __popc(__ldcg(b.scratch_ballots + t));
    scan[threadIdx.x] = cnt;
    __syncthreads();
    int off = 0, total = 0;
    for (int i = 0; i < (int)blockDim.x; ++i) {
        const int c = scan[i];
        off += (i < (int)threadIdx.x) ? c : 0;
        total += c;
    }
    double acc[HB_NUM_REWARDS];
#pragma unroll
    for (int k = 0; k < HB_NUM_REWARDS; ++k) acc[k] = 0.0;
    for (int t = t_lo; t < t_hi; ++t) {
        unsigned m = __ldcg(b.scratch_ballots + t);
        if (!m) continue;
#pragma unroll
        for (int k = 0; k < HB_NUM_REWARDS; ++k) acc[k] += (double)__ldcg(b.scratch_partials + (size_t)t * HB_NUM_REWARDS + k);
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            b.reset_env_ids[off++] = t * TILE + bit;
        }
    }
#pragma unroll
    for (int k = 0; k < HB_NUM_REWARDS; ++k) red[k * blockDim.x + threadIdx.x] = acc[k];
    __syncthreads();
    // extras["episode"] is only refreshed on steps with >= 1 reset (quirk 4): otherwise the previous
    // values are carried into this step's slot.
    if (threadIdx.x < HB_NUM_REWARDS) {
        const int k = threadIdx.x;
        if (total > 0) {
            double v = 0.0;
            for (int i = 0; i < (int)blockDim.x; ++i) v += red[k * blockDim.x + i];
            b.episode_means[k] = (float)(v / (double)total) / p.max_episode_length_s;
        } else if (b.episode_means_prev && b.episode_means_prev != b.episode_means) {
            b.episode_means[k] = b.episode_means_prev[k];
        }
    }
    if (threadIdx.x == 0) {
        *b.reset_count = total;
        if (host_count) *host_count = total;
        *b.scratch_ticket = 0u;          // re-arm for the next launch
    }
}

// ------------------------------------------------------------------------------------------
// a8 (stacking): new[:, 0:(S-1)*F] = reset ? 0 : prev[:, F:S*F]  — a flat copy shifted by one
// frame with a hole of F floats per row (the newest frame, written by post_physics_kernel).
// Destination vectors are 16-byte aligned; the source is 4*F bytes further on, which is only
// 4-byte aligned for F = 41, so each thread loads the aligned vector below its source window
// and takes the missing float from its neighbour lane (shuffle); F % 4 picks the rotation.
// ------------------------------------------------------------------------------------------
template <int UNROLL>
__device__ __forceinline__ void
stack_shift_body(const float *__restrict__ prev, float *__restrict__ next, const uint8_t *__restrict__ reset_buf,
                 long long total, int row, int frame, const uint8_t *__restrict__ latch_src,
                 uint8_t *__restrict__ latch_dst, const int32_t *__restrict__ reset_count, int num_envs,
                 const unsigned blk, const unsigned nblk) {
    const int lane = threadIdx.x & 31;
    const int keep = row - frame;                       // floats of a row that are carried over
    const int rot = frame & 3;                          // source misalignment in floats
    const int fvec = frame >> 2;                        // whole vectors of shift
    const long long warp_base = ((long long)blk * blockDim.x + threadIdx.x - lane) * UNROLL;
    // extras["time_outs"] latch (legged_robot.py:208-209 only runs when >= 1 env was reset)
    if (latch_dst && *reset_count > 0) {
        for (long long i = (long long)blk * blockDim.x + threadIdx.x; i < num_envs; i += (long long)nblk * blockDim.x)
            latch_dst[i] = latch_src[i];
    }
    const float4 *p4 = reinterpret_cast<const float4 *>(prev);
    float4 *n4 = reinterpret_cast<float4 *>(next);
    const long long total_vec = total >> 2;             // whole vectors; a ragged tail (N*row % 4) goes scalar
    const long long tail_vec = (total + 3) >> 2;
    float4 v[UNROLL];
    float nx31[UNROLL][3];
    // all loads first (UNROLL independent 16-byte requests in flight per thread)
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const long long i = warp_base + (long long)u * 32 + lane;      // destination vector index
        const long long s = i + fvec;                                   // aligned source vector below the window
        if (s < total_vec) {
            v[u] = hb::ld_stream4(p4 + s);
        } else {                                                        // ragged last vector of the buffer
            const long long e = s * 4;
            v[u].x = (e < total) ? __ldg(prev + e) : 0.0f;
            v[u].y = (e + 1 < total) ? __ldg(prev + e + 1) : 0.0f;
            v[u].z = (e + 2 < total) ? __ldg(prev + e + 2) : 0.0f;
            v[u].w = 0.0f;
        }
        nx31[u][0] = nx31[u][1] = nx31[u][2] = 0.0f;
        if (lane == 31 && rot != 0) {                                   // no neighbour lane: fetch the spill-over
            const long long e = (s + 1) * 4;
            if (e < total) nx31[u][0] = __ldg(prev + e);
            if (rot > 1 && e + 1 < total) nx31[u][1] = __ldg(prev + e + 1);
            if (rot > 2 && e + 2 < total) nx31[u][2] = __ldg(prev + e + 2);
        }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const long long i = warp_base + (long long)u * 32 + lane;
        float nx0 = __shfl_down_sync(0xffffffffu, v[u].x, 1);           // first floats of vector s+1
        float nx1 = __shfl_down_sync(0xffffffffu, v[u].y, 1);
        float nx2 = __shfl_down_sync(0xffffffffu, v[u].z, 1);
        if (lane == 31) nx0 = nx31[u][0], nx1 = nx31[u][1], nx2 = nx31[u][2];
        if (i >= tail_vec) continue;
        float4 o;
        if (rot == 0) o = v[u];
        else if (rot == 1) o = make_float4(v[u].y, v[u].z, v[u].w, nx0);
        else if (rot == 2) o = make_float4(v[u].z, v[u].w, nx0, nx1);
        else o = make_float4(v[u].w, nx0, nx1, nx2);
        const long long e0 = i * 4;                      // first destination float
        const int r0 = (int)(e0 / row);
        const int c0 = (int)(e0 - (long long)r0 * row);
        if (c0 + 3 < keep && i < total_vec) {            // whole vector inside the carried part of one row
            if (reset_buf[r0]) o = make_float4(0.f, 0.f, 0.f, 0.f);
            hb::st_stream4(n4 + i, o);
        } else {                                         // touches the newest-frame hole or a row boundary
            const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int r = r0, c = c0 + k;
                if (c >= row) c -= row, r += 1;
                if (c < keep && e0 + k < total) next[e0 + k] = reset_buf[r] ? 0.0f : ov[k];
            }
        }
    }
}

template <int UNROLL>
__global__ void __launch_bounds__(256)
stack_shift_kernel(const float *__restrict__ prev, float *__restrict__ next, const uint8_t *__restrict__ reset_buf,
                   long long total, int row, int frame, const uint8_t *__restrict__ latch_src,
                   uint8_t *__restrict__ latch_dst, const int32_t *__restrict__ reset_count, int num_envs) {
    stack_shift_body<UNROLL>(prev, next, reset_buf, total, row, frame, latch_src, latch_dst, reset_count, num_envs,
                             blockIdx.x, gridDim.x);
}

// actor and critic histories in one launch: blocks [0, blocks_a) shift buffer a, the rest buffer b
template <int UNROLL>
__global__ void __launch_bounds__(256)
stack_shift_pair_kernel(const float *__restrict__ prev_a, float *__restrict__ next_a, long long total_a, int row_a,
                        int frame_a, unsigned blocks_a, const float *__restrict__ prev_b, float *__restrict__ next_b,
                        long long total_b, int row_b, int frame_b, const uint8_t *__restrict__ reset_buf,
                        const uint8_t *__restrict__ latch_src, uint8_t *__restrict__ latch_dst,
                        const int32_t *__restrict__ reset_count, int num_envs) {
    if (blockIdx.x < blocks_a)
        stack_shift_body<UNROLL>(prev_a, next_a, reset_buf, total_a, row_a, frame_a, latch_src, latch_dst, reset_count,
                                 num_envs, blockIdx.x, blocks_a);
    else
        stack_shift_body<UNROLL>(prev_b, next_b, reset_buf, total_b, row_b, frame_b, nullptr, nullptr, nullptr, num_envs,
                                 blockIdx.x - blocks_a, gridDim.x - blocks_a);
}

// generic fallback (rows not a multiple of 4 floats in total, or unaligned buffers)
__global__ void __launch_bounds__(256)
stack_shift_scalar_kernel(const float *__restrict__ prev, float *__restrict__ next,
                          const uint8_t *__restrict__ reset_buf, long long total, int row, int frame,
                          const uint8_t *__restrict__ latch_src, uint8_t *__restrict__ latch_dst,
                          const int32_t *__restrict__ reset_count, int num_envs) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (latch_dst && *reset_count > 0 && i < num_envs) latch_dst[i] = latch_src[i];
    if (i >= total) return;
    const int r = (int)(i / row), c = (int)(i - (long long)r * row);
    if (c < row - frame) next[i] = reset_buf[r] ? 0.0f : prev[i + frame];
}

int g_use_bulk = 1;

}  // namespace

// ==============================================================================================
// C ABI
// ==============================================================================================
extern "C" {

int hb_set_option(const char *name, int value) {
    if (name && !strcmp(name, "env_bulk_staging")) {
        g_use_bulk = value;
        return HB_OK;
    }
    hb::set_error("hb_set_option: unknown option '%s'", name ? name : "(null)");
    return HB_ERR_BAD_ARG;
}

static int check_params(const hb_env_params *p, const hb_env_buffers *buf, const char *who) {
    HB_REQUIRE(p && buf, "%s: null params/buffers", who);
    HB_REQUIRE(p->abi_version == HB_ABI_VERSION, "%s: ABI version %d != %d", who, p->abi_version, HB_ABI_VERSION);
    HB_REQUIRE(p->num_envs > 0, "%s: num_envs must be positive", who);
    HB_REQUIRE(p->resample_interval > 0, "%s: resample_interval must be positive", who);
    HB_REQUIRE(p->num_dof == NDOF, "%s: only the 10-DOF hector layout is built (num_dof=%d)", who, p->num_dof);
    return HB_OK;
}

int hb_env_action_prologue(const hb_env_params *p, const hb_env_buffers *buf, const float *actions_in,
                           const hb_env_noise *noise, void *stream) {
    if (int rc = check_params(p, buf, "hb_env_action_prologue")) return rc;
    HB_REQUIRE(actions_in && buf->actions, "hb_env_action_prologue: null actions");
    const int total = p->num_envs * p->num_dof;
    action_prologue_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        actions_in, buf->actions, noise ? noise->u_delay : nullptr, noise ? noise->z_action : nullptr, total,
        p->clip_actions, p->action_delay, p->action_noise);
    HB_CHECK_LAUNCH("action_prologue_kernel");
    return HB_OK;
}

int hb_env_compute_torques(const hb_env_params *p, const hb_env_buffers *buf, void *stream) {
    if (int rc = check_params(p, buf, "hb_env_compute_torques")) return rc;
    HB_REQUIRE(buf->dof_state && buf->actions && buf->p_gains && buf->d_gains && buf->torques,
               "hb_env_compute_torques: null buffer");
    PdConsts c;
    for (int j = 0; j < HB_MAX_DOF; ++j) c.q0[j] = p->default_dof_pos[j], c.lim[j] = p->torque_limits[j];
    const int total = p->num_envs * p->num_dof;
    const bool vec = (p->num_dof % 2 == 0) && hb::aligned16(buf->dof_state) &&
                     ((reinterpret_cast<uintptr_t>(buf->actions) | reinterpret_cast<uintptr_t>(buf->p_gains) |
                       reinterpret_cast<uintptr_t>(buf->d_gains) | reinterpret_cast<uintptr_t>(buf->torques)) & 7u) == 0;
    if (vec) {
        const int pairs = total / 2;
        pd_torque_kernel<<<(pairs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4 *>(buf->dof_state), reinterpret_cast<const float2 *>(buf->actions),
            reinterpret_cast<const float2 *>(buf->p_gains), reinterpret_cast<const float2 *>(buf->d_gains),
            reinterpret_cast<float2 *>(buf->torques), pairs, p->num_dof, p->action_scale, c);
    } else {
        pd_torque_scalar_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
            buf->dof_state, buf->actions, buf->p_gains, buf->d_gains, buf->torques, total, p->num_dof,
            p->action_scale, c);
    }
    HB_CHECK_LAUNCH("pd_torque_kernel");
    return HB_OK;
}

int hb_env_post_physics(const hb_env_params *p, const hb_env_buffers *buf, const hb_env_noise *noise,
                        float *obs_new, float *priv_new, int32_t stages, int32_t *host_count, void *stream) {
    if (int rc = check_params(p, buf, "hb_env_post_physics")) return rc;
    HB_REQUIRE(noise && obs_new && priv_new, "hb_env_post_physics: null noise/obs pointers");
    HB_REQUIRE(p->num_single_obs == OBS && p->num_single_priv == PRIV,
               "hb_env_post_physics: frame sizes %d/%d do not match the hector layout %d/%d", p->num_single_obs,
               p->num_single_priv, OBS, PRIV);
    HB_REQUIRE(p->num_bodies > 0 && p->num_bodies <= 32, "hb_env_post_physics: num_bodies out of range");
    HB_REQUIRE(p->n_term >= 0 && p->n_term <= HB_MAX_CONTACT_BODIES && p->n_pen >= 0 &&
                   p->n_pen <= HB_MAX_CONTACT_BODIES, "hb_env_post_physics: too many contact bodies");
    HB_REQUIRE(noise->u_reset, "hb_env_post_physics: u_reset is required (any env may reset)");
    HB_REQUIRE(buf->scratch_ballots && buf->scratch_partials && buf->scratch_ticket && buf->reset_env_ids &&
                   buf->reset_count && buf->episode_means, "hb_env_post_physics: null scratch/result buffers");
    const bool with_noise = p->add_noise && noise->z_obs;
    const TileLayout L = make_layout(p->num_bodies, with_noise);
    const size_t smem = (size_t)L.total * sizeof(float);
    const int tiles = (p->num_envs + TILE - 1) / TILE;
    // bulk staging needs 16-byte aligned slabs: base pointers aligned and 32-env tiles (128-byte multiples)
    bool bulk = g_use_bulk != 0;
    const void *slabs[] = {buf->root_states, buf->dof_state, buf->contact_forces, buf->actions, buf->last_actions,
                           buf->last_last_actions, buf->last_dof_vel, buf->torques, buf->last_root_vel,
                           buf->commands, with_noise ? noise->z_obs : buf->commands};
    bool all_aligned = true;
    for (const void *s : slabs) all_aligned = all_aligned && hb::aligned16(s);
    HB_REQUIRE(all_aligned, "hb_env_post_physics: state tensors must be 16-byte aligned");
    static bool attr_set[2] = {false, false};
    if (bulk) {
        if (!attr_set[1]) {
            HB_CUDA(cudaFuncSetAttribute(post_physics_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            attr_set[1] = true;
        }
        post_physics_kernel<true><<<tiles, 4 * TILE, smem, (cudaStream_t)stream>>>(*p, *buf, *noise, obs_new, priv_new,
                                                                               stages, host_count);
    } else {
        if (!attr_set[0]) {
            HB_CUDA(cudaFuncSetAttribute(post_physics_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            attr_set[0] = true;
        }
        post_physics_kernel<false><<<tiles, 4 * TILE, smem, (cudaStream_t)stream>>>(*p, *buf, *noise, obs_new, priv_new,
                                                                                stages, host_count);
    }
    HB_CHECK_LAUNCH("post_physics_kernel");
    return HB_OK;
}

static int launch_stack(const float *prev, float *next, const uint8_t *reset_buf, int n, int row, int frame,
                        const uint8_t *latch_src, uint8_t *latch_dst, const int32_t *reset_count, cudaStream_t st) {
    const long long total = (long long)n * row;
    if (hb::aligned16(prev) && hb::aligned16(next)) {
        constexpr int UNROLL = 4;
        const long long vecs = (total + 3) / 4;
        const long long threads = (vecs + UNROLL - 1) / UNROLL;
        const int blocks = (int)((threads + 255) / 256);
        stack_shift_kernel<UNROLL><<<blocks, 256, 0, st>>>(prev, next, reset_buf, total, row, frame, latch_src,
                                                            latch_dst, reset_count, n);
    } else {
        stack_shift_scalar_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(prev, next, reset_buf, total, row, frame,
                                                                               latch_src, latch_dst, reset_count, n);
    }
    return 0;
}

int hb_stack_shift(const float *prev, float *next, const uint8_t *reset_buf, int32_t num_envs, int32_t row,
                   int32_t frame, void *stream) {
    HB_REQUIRE(prev && next && reset_buf && prev != next, "hb_stack_shift: null or aliasing buffers");
    HB_REQUIRE(num_envs > 0 && frame > 0 && row > frame, "hb_stack_shift: bad shape");
    launch_stack(prev, next, reset_buf, num_envs, row, frame, nullptr, nullptr, nullptr, (cudaStream_t)stream);
    HB_CHECK_LAUNCH("stack_shift_kernel");
    return HB_OK;
}

int hb_env_stack_observations(const hb_env_params *p, const hb_env_buffers *buf, const float *obs_prev,
                              const float *priv_prev, float *obs_new, float *priv_new, void *stream) {
    if (int rc = check_params(p, buf, "hb_env_stack_observations")) return rc;
    HB_REQUIRE(obs_prev && priv_prev && obs_new && priv_new && buf->reset_buf, "hb_env_stack_observations: null buffer");
    HB_REQUIRE(obs_prev != obs_new && priv_prev != priv_new, "hb_env_stack_observations: prev and new must not alias");
    cudaStream_t st = (cudaStream_t)stream;
    if (hb::aligned16(obs_prev) && hb::aligned16(obs_new) && hb::aligned16(priv_prev) && hb::aligned16(priv_new)) {
        constexpr int UNROLL = 4;
        const int row_a = p->frame_stack * p->num_single_obs, row_b = p->c_frame_stack * p->num_single_priv;
        const long long total_a = (long long)p->num_envs * row_a, total_b = (long long)p->num_envs * row_b;
        const unsigned blocks_a = (unsigned)(((total_a + 3) / 4 + UNROLL * 256 - 1) / (UNROLL * 256));
        const unsigned blocks_b = (unsigned)(((total_b + 3) / 4 + UNROLL * 256 - 1) / (UNROLL * 256));
        stack_shift_pair_kernel<UNROLL><<<blocks_a + blocks_b, 256, 0, st>>>(
            obs_prev, obs_new, total_a, row_a, p->num_single_obs, blocks_a, priv_prev, priv_new, total_b, row_b,
            p->num_single_priv, buf->reset_buf, buf->time_out_buf, buf->time_outs_latched, buf->reset_count,
            p->num_envs);
        HB_CHECK_LAUNCH("stack_shift_pair_kernel");
        return HB_OK;
    }
    launch_stack(obs_prev, obs_new, buf->reset_buf, p->num_envs, p->frame_stack * p->num_single_obs, p->num_single_obs,
                 buf->time_out_buf, buf->time_outs_latched, buf->reset_count, st);
    HB_CHECK_LAUNCH("stack_shift_kernel(obs)");
    launch_stack(priv_prev, priv_new, buf->reset_buf, p->num_envs, p->c_frame_stack * p->num_single_priv,
                 p->num_single_priv, nullptr, nullptr, nullptr, st);
    HB_CHECK_LAUNCH("stack_shift_kernel(priv)");
    return HB_OK;
}

}  // extern "C"
