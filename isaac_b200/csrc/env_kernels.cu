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

// get_euler_xyz_tensor (envs/base/legged_robot.py:50-55) over isaacgym get_euler_xyz
__device__ __forceinline__ Vec3 euler_xyz_wrapped(const float q[4]) {
    const float x = q[0], y = q[1], z = q[2], w = q[3];
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
    roll = py_mod(roll, TWO_PI_F), pitch = py_mod(pitch, TWO_PI_F), yaw = py_mod(yaw, TWO_PI_F);
    if (roll > PI_F) roll -= TWO_PI_F;
    if (pitch > PI_F) pitch -= TWO_PI_F;
    if (yaw > PI_F) yaw -= TWO_PI_F;
    return {roll, pitch, yaw};
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
// a3-a9: post_physics_step, one warp per 32-env tile, one lane per env
// ------------------------------------------------------------------------------------------
struct TileLayout {          // float offsets into the warp's shared-memory tile
    int root, dof, contact, actions, last_actions, last_last_actions, last_dof_vel, torques, last_root_vel,
        commands, z_obs, total;
};

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
    const int out = TILE * (OBS + PRIV);      // the frames are staged over the consumed inputs
    L.total = o > out ? o : out;
    return L;
}

__device__ __forceinline__ void stage_ldg(float *dst, const float *__restrict__ src, int count, int lane, bool vec) {
    if (vec) {
        const float4 *s4 = reinterpret_cast<const float4 *>(src);
        float4 *d4 = reinterpret_cast<float4 *>(dst);
        for (int i = lane; i < count / 4; i += 32) d4[i] = __ldg(s4 + i);
    } else {
        for (int i = lane; i < count; i += 32) dst[i] = __ldg(src + i);
    }
}

template <bool kBulk>
__global__ void __launch_bounds__(TILE)
post_physics_kernel(const __grid_constant__ hb_env_params p, const __grid_constant__ hb_env_buffers b,
                    const __grid_constant__ hb_env_noise nz, float *__restrict__ obs_new,
                    float *__restrict__ priv_new, int stages, int32_t *host_count) {
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int s_is_last;
    const int lane = threadIdx.x;
    const int N = p.num_envs;
    const int env0 = blockIdx.x * TILE;
    const int nv = min(TILE, N - env0);
    const int env = env0 + lane;
    const bool valid = lane < nv;
    const bool do_step = stages & HB_STAGE_STEP;
    const bool reset_all = stages & HB_STAGE_RESET_ALL;
    const bool with_noise = p.add_noise && nz.z_obs != nullptr;
    const TileLayout L = make_layout(p.num_bodies, with_noise);
    const int crow = p.num_bodies * 3;

    // ---------------- stage the tile's input slabs into shared memory ----------------
    const bool full = (nv == TILE);
    if (kBulk && full) {
        if (lane == 0) {
            hb::mbar_init(&bar, 1);
            hb::fence_mbar_init();
        }
        __syncwarp();
        if (lane == 0) {
            uint32_t bytes = TILE * 4u * (13 + 2 * NDOF + crow + 5 * NDOF + 6 + 4 + (with_noise ? OBS : 0));
            hb::mbar_expect_tx(&bar, bytes);
            const size_t e = env0;
            hb::bulk_g2s(sm + L.root, b.root_states + e * 13, TILE * 13 * 4, &bar);
            hb::bulk_g2s(sm + L.dof, b.dof_state + e * NDOF * 2, TILE * NDOF * 2 * 4, &bar);
            hb::bulk_g2s(sm + L.contact, b.contact_forces + e * crow, TILE * crow * 4, &bar);
            hb::bulk_g2s(sm + L.actions, b.actions + e * NDOF, TILE * NDOF * 4, &bar);
            hb::bulk_g2s(sm + L.last_actions, b.last_actions + e * NDOF, TILE * NDOF * 4, &bar);
            hb::bulk_g2s(sm + L.last_last_actions, b.last_last_actions + e * NDOF, TILE * NDOF * 4, &bar);
            hb::bulk_g2s(sm + L.last_dof_vel, b.last_dof_vel + e * NDOF, TILE * NDOF * 4, &bar);
            hb::bulk_g2s(sm + L.torques, b.torques + e * NDOF, TILE * NDOF * 4, &bar);
            hb::bulk_g2s(sm + L.last_root_vel, b.last_root_vel + e * 6, TILE * 6 * 4, &bar);
            hb::bulk_g2s(sm + L.commands, b.commands + e * 4, TILE * 4 * 4, &bar);
            if (with_noise) hb::bulk_g2s(sm + L.z_obs, nz.z_obs + e * OBS, TILE * OBS * 4, &bar);
        }
    } else {
        const size_t e = env0;
        stage_ldg(sm + L.root, b.root_states + e * 13, nv * 13, lane, full);
        stage_ldg(sm + L.dof, b.dof_state + e * NDOF * 2, nv * NDOF * 2, lane, full);
        stage_ldg(sm + L.contact, b.contact_forces + e * crow, nv * crow, lane, full);
        stage_ldg(sm + L.actions, b.actions + e * NDOF, nv * NDOF, lane, full);
        stage_ldg(sm + L.last_actions, b.last_actions + e * NDOF, nv * NDOF, lane, full);
        stage_ldg(sm + L.last_last_actions, b.last_last_actions + e * NDOF, nv * NDOF, lane, full);
        stage_ldg(sm + L.last_dof_vel, b.last_dof_vel + e * NDOF, nv * NDOF, lane, full);
        stage_ldg(sm + L.torques, b.torques + e * NDOF, nv * NDOF, lane, full);
        stage_ldg(sm + L.last_root_vel, b.last_root_vel + e * 6, nv * 6, lane, full);
        stage_ldg(sm + L.commands, b.commands + e * 4, nv * 4, lane, full);
        if (with_noise) stage_ldg(sm + L.z_obs, nz.z_obs + e * OBS, nv * OBS, lane, full);
    }

    // ---------------- per-lane direct loads that overlap the staging ----------------
    // rigid body rows: feet pos(0:3)+lin vel(7:10), knees xy (strided 52-byte rows; 16 of 143 floats)
    float foot_pos[2][3], foot_vel[2][3], knee_xy[2][2];
    float sums[HB_NUM_REWARDS];
    long long ep_len = 0;
    float air[2], fh[2], lz[2];
    bool last_ct[2];
    float push_f[2], push_t[3];
    if (valid) {
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
        ep_len = b.episode_length_buf[env];
#pragma unroll
        for (int k = 0; k < HB_NUM_REWARDS; ++k) sums[k] = b.episode_sums[(size_t)k * N + env];
        push_f[0] = b.rand_push_force[env * 3], push_f[1] = b.rand_push_force[env * 3 + 1];
#pragma unroll
        for (int k = 0; k < 3; ++k) push_t[k] = b.rand_push_torque[env * 3 + k];
    }

    if (kBulk && full) {
        hb::mbar_wait(&bar, 0);
    }
    __syncwarp();

    // ---------------- unpack this lane's rows ----------------
    const int ln = valid ? lane : 0;
    float root[13], q[NDOF], qd[NDOF], act[NDOF], lact[NDOF], llact[NDOF], ldv[NDOF], tau[NDOF], lrv[6], cmd[4];
#pragma unroll
    for (int k = 0; k < 13; ++k) root[k] = sm[L.root + ln * 13 + k];
#pragma unroll
    for (int j = 0; j < NDOF; ++j) {
        q[j] = sm[L.dof + ln * NDOF * 2 + 2 * j];
        qd[j] = sm[L.dof + ln * NDOF * 2 + 2 * j + 1];
        act[j] = sm[L.actions + ln * NDOF + j];
        lact[j] = sm[L.last_actions + ln * NDOF + j];
        llact[j] = sm[L.last_last_actions + ln * NDOF + j];
        ldv[j] = sm[L.last_dof_vel + ln * NDOF + j];
        tau[j] = sm[L.torques + ln * NDOF + j];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) lrv[k] = sm[L.last_root_vel + ln * 6 + k];
#pragma unroll
    for (int k = 0; k < 4; ++k) cmd[k] = sm[L.commands + ln * 4 + k];
    const float *cf = sm + L.contact + ln * crow;
    float foot_f[2][3];
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int k = 0; k < 3; ++k) foot_f[f][k] = cf[p.feet[f] * 3 + k];
    float term_norm[HB_MAX_CONTACT_BODIES], pen_norm[HB_MAX_CONTACT_BODIES];
#pragma unroll
    for (int i = 0; i < HB_MAX_CONTACT_BODIES; ++i) {
        term_norm[i] = pen_norm[i] = 0.0f;
        if (i < p.n_term) {
            const float *f3 = cf + p.term_bodies[i] * 3;
            term_norm[i] = sqrtf((f3[0] * f3[0] + f3[1] * f3[1]) + f3[2] * f3[2]);
        }
        if (i < p.n_pen) {
            const float *f3 = cf + p.pen_bodies[i] * 3;
            pen_norm[i] = sqrtf((f3[0] * f3[0] + f3[1] * f3[1]) + f3[2] * f3[2]);
        }
    }
    float zobs[OBS];
#pragma unroll
    for (int k = 0; k < OBS; ++k) zobs[k] = with_noise ? sm[L.z_obs + ln * OBS + k] : 0.0f;
    __syncwarp();      // inputs consumed: the tile memory is reused for the output frames below

    Vec3 lin = {0.f, 0.f, 0.f}, ang = {0.f, 0.f, 0.f}, grav = {0.f, 0.f, 0.f}, eul = {0.f, 0.f, 0.f};
    bool reset = false, time_out = false;
    float rew = 0.0f;
    const float dt = p.dt;

    if (do_step) {
        // -------- legged_robot.py:127-135 --------
        ep_len += 1;
        const float *quat = root + 3;
        lin = quat_rotate_inverse(quat, {root[7], root[8], root[9]});
        ang = quat_rotate_inverse(quat, {root[10], root[11], root[12]});
        grav = quat_rotate_inverse(quat, {0.0f, 0.0f, -1.0f});
        eul = euler_xyz_wrapped(quat);

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
            float e = py_mod(cmd[3] - heading, TWO_PI_F);          // wrap_to_pi, utils/math.py:46-49
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
            b.root_states[(size_t)env * 13 + 7] = root[7];
            b.root_states[(size_t)env * 13 + 8] = root[8];
#pragma unroll
            for (int k = 0; k < 3; ++k) b.root_states[(size_t)env * 13 + 10 + k] = root[10 + k];
            b.rand_push_force[env * 3] = push_f[0], b.rand_push_force[env * 3 + 1] = push_f[1];
#pragma unroll
            for (int k = 0; k < 3; ++k) b.rand_push_torque[env * 3 + k] = push_t[k];
        }

        // -------- check_termination, legged_robot.py:155-160 --------
#pragma unroll
        for (int i = 0; i < HB_MAX_CONTACT_BODIES; ++i) reset |= (i < p.n_term) && (term_norm[i] > 1.0f);
        time_out = ep_len > (long long)p.max_episode_length;
        reset |= time_out;

        // -------- gait phase (hector_env.py:70-88) and contacts --------
        const float phase = ((float)ep_len * dt) / p.cycle_time;
        const float s = sinf(TWO_PI_F * phase);
        float st[2] = {(s >= 0.0f) ? 1.0f : 0.0f, (s < 0.0f) ? 1.0f : 0.0f};
        if (fabsf(s) < 0.1f) st[0] = st[1] = 1.0f;
        const bool ct[2] = {foot_f[0][2] > 5.0f, foot_f[1][2] > 5.0f};

        // -------- compute_reward (legged_robot.py:216-234): alphabetical accumulation --------
        float term[HB_NUM_REWARDS];
        {   // action_smoothness, hector_env.py:529-539
            float t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll
            for (int j = 0; j < NDOF; ++j) {
                const float d1 = lact[j] - act[j];
                t1 += d1 * d1;
                const float d2 = (act[j] + llact[j]) - 2.0f * lact[j];
                t2 += d2 * d2;
                t3 += fabsf(act[j]);
            }
            term[HB_R_ACTION_SMOOTHNESS] = (t1 + t2) + 0.05f * t3;
        }
        {   // base_acc, :385-392
            float ss = 0.f;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const float d = lrv[k] - root[7 + k];
                ss += d * d;
            }
            term[HB_R_BASE_ACC] = expf(-sqrtf(ss) * 3.0f);
        }
        {   // base_height, :369-383
            const float ground = (foot_pos[0][2] * st[0] + foot_pos[1][2] * st[1]) / (st[0] + st[1]);
            const float h = root[2] - (ground - 0.05f);
            term[HB_R_BASE_HEIGHT] = expf(-fabsf(h - p.base_height_target) * 100.0f);
        }
        {   // collision, :522-527
            float c = 0.f;
#pragma unroll
            for (int i = 0; i < HB_MAX_CONTACT_BODIES; ++i) c += ((i < p.n_pen) && (pen_norm[i] > 0.1f)) ? 1.0f : 0.0f;
            term[HB_R_COLLISION] = c;
        }
        {   // default_joint_pos, :357-367
            float d[NDOF], ss = 0.f;
#pragma unroll
            for (int j = 0; j < NDOF; ++j) d[j] = q[j] - p.default_dof_pos[j], ss += d[j] * d[j];
            float yr = sqrtf(d[0] * d[0] + d[1] * d[1]) + sqrtf(d[5] * d[5] + d[6] * d[6]);
            yr = clampf(yr - 0.1f, 0.0f, 50.0f);
            term[HB_R_DEFAULT_JOINT_POS] = expf(-yr * 100.0f) - 0.01f * sqrtf(ss);
        }
        {   // dof_acc :515-520, dof_vel :508-513, torques :501-506
            float acc = 0.f, vel = 0.f, tq = 0.f;
#pragma unroll
            for (int j = 0; j < NDOF; ++j) {
                const float a = (ldv[j] - qd[j]) / dt;
                acc += a * a;
                vel += qd[j] * qd[j];
                tq += tau[j] * tau[j];
            }
            term[HB_R_DOF_ACC] = acc, term[HB_R_DOF_VEL] = vel, term[HB_R_TORQUES] = tq;
        }
        {   // feet_air_time, :315-329 (stateful)
            float r = 0.f;
#pragma unroll
            for (int f = 0; f < 2; ++f) {
                const bool filt = ct[f] || (st[f] != 0.0f) || last_ct[f];
                last_ct[f] = ct[f];
                const bool first = (air[f] > 0.0f) && filt;
                air[f] += dt;
                r += clampf(air[f], 0.0f, 0.5f) * (first ? 1.0f : 0.0f);
                air[f] *= filt ? 0.0f : 1.0f;
            }
            term[HB_R_FEET_AIR_TIME] = r;
        }
        {   // feet_clearance, :445-466 (stateful)
            float r = 0.f;
#pragma unroll
            for (int f = 0; f < 2; ++f) {
                const float z = foot_pos[f][2] - 0.05f;
                fh[f] += z - lz[f];
                lz[f] = z;
                const float hit = (fabsf(fh[f] - p.target_feet_height) < 0.01f) ? 1.0f : 0.0f;
                r += hit * (1.0f - st[f]);
                fh[f] *= ct[f] ? 0.0f : 1.0f;
            }
            term[HB_R_FEET_CLEARANCE] = r;
        }
        {   // feet_contact_forces :350-355, feet_contact_number :331-339, foot_slip :303-313
            float cfz = 0.f, num = 0.f, slip = 0.f;
#pragma unroll
            for (int f = 0; f < 2; ++f) {
                const float nf = sqrtf((foot_f[f][0] * foot_f[f][0] + foot_f[f][1] * foot_f[f][1]) + foot_f[f][2] * foot_f[f][2]);
                cfz += clampf(nf - p.max_contact_force, 0.0f, 400.0f);
                num += (ct[f] == (st[f] != 0.0f)) ? 1.0f : -0.3f;
                const float sp = sqrtf(sqrtf(foot_vel[f][0] * foot_vel[f][0] + foot_vel[f][1] * foot_vel[f][1]));
                slip += sp * (ct[f] ? 1.0f : 0.0f);
            }
            term[HB_R_FEET_CONTACT_FORCES] = cfz;
            term[HB_R_FEET_CONTACT_NUMBER] = num / 2.0f;
            term[HB_R_FOOT_SLIP] = slip;
        }
        {   // feet_distance :277-287, knee_distance :290-300
            float dx = foot_pos[0][0] - foot_pos[1][0], dy = foot_pos[0][1] - foot_pos[1][1];
            float d = sqrtf(dx * dx + dy * dy);
            float dmin = clampf(d - p.min_dist, -0.5f, 0.0f), dmax = clampf(d - p.max_dist, 0.0f, 0.5f);
            term[HB_R_FEET_DISTANCE] = (expf(-fabsf(dmin) * 100.0f) + expf(-fabsf(dmax) * 100.0f)) / 2.0f;
            dx = knee_xy[0][0] - knee_xy[1][0], dy = knee_xy[0][1] - knee_xy[1][1];
            d = sqrtf(dx * dx + dy * dy);
            dmin = clampf(d - p.min_dist, -0.5f, 0.0f), dmax = clampf(d - p.max_dist / 2.0f, 0.0f, 0.5f);
            term[HB_R_KNEE_DISTANCE] = (expf(-fabsf(dmin) * 100.0f) + expf(-fabsf(dmax) * 100.0f)) / 2.0f;
        }
        {   // orientation, :341-348
            const float a = expf(-(fabsf(eul.x) + fabsf(eul.y)) * 10.0f);
            const float g = expf(-sqrtf(grav.x * grav.x + grav.y * grav.y) * 20.0f);
            term[HB_R_ORIENTATION] = (a + g) / 2.0f;
        }
        {   // tracking_ang_vel :435-443, tracking_lin_vel :426-433
            const float ea = (cmd[2] - ang.z) * (cmd[2] - ang.z);
            term[HB_R_TRACKING_ANG_VEL] = expf(-ea * p.tracking_sigma);
            const float ex = cmd[0] - lin.x, ey = cmd[1] - lin.y;
            term[HB_R_TRACKING_LIN_VEL] = expf(-(ex * ex + ey * ey) * p.tracking_sigma);
        }
#pragma unroll
        for (int k = 0; k < HB_NUM_REWARDS; ++k) {
            if (p.reward_scale[k] != 0.0f) {
                const float r = term[k] * p.reward_scale[k];
                rew += r;
                sums[k] += r;
            }
        }
        if (p.only_positive_rewards) rew = fmaxf(rew, 0.0f);
    } else if (stages & HB_STAGE_DERIVE) {
        // _init_buffers (legged_robot.py:452,477-479): derived quantities of the initial state
        const float *quat = root + 3;
        lin = quat_rotate_inverse(quat, {root[7], root[8], root[9]});
        ang = quat_rotate_inverse(quat, {root[10], root[11], root[12]});
        grav = quat_rotate_inverse(quat, {0.0f, 0.0f, -1.0f});
        eul = euler_xyz_wrapped(quat);
    } else if (valid) {
        // reset / observation-only pass: derived quantities keep their stored values
        lin = {b.base_lin_vel[env * 3], b.base_lin_vel[env * 3 + 1], b.base_lin_vel[env * 3 + 2]};
        ang = {b.base_ang_vel[env * 3], b.base_ang_vel[env * 3 + 1], b.base_ang_vel[env * 3 + 2]};
        grav = {b.projected_gravity[env * 3], b.projected_gravity[env * 3 + 1], b.projected_gravity[env * 3 + 2]};
        eul = euler_xyz_wrapped(root + 3);
    }
    if (reset_all) reset = true;
    if ((stages & HB_STAGE_RESET_MASK) && valid) reset = reset || (b.reset_buf[env] != 0);   // reset_idx(env_ids)
    reset = reset && valid;

    // ---------------- reset_idx (legged_robot.py:162-214,358-396) ----------------
    const unsigned ballot = __ballot_sync(0xffffffffu, reset);
    float my_sums[HB_NUM_REWARDS];
#pragma unroll
    for (int k = 0; k < HB_NUM_REWARDS; ++k) my_sums[k] = reset ? sums[k] : 0.0f;
    if (reset) {
        const float *u = nz.u_reset + (size_t)env * 15;
#pragma unroll
        for (int j = 0; j < NDOF; ++j) {
            q[j] = p.default_dof_pos[j] + (p.reset_dof_span * u[j] + p.reset_dof_lo);
            qd[j] = 0.0f;
            b.dof_state[((size_t)env * NDOF + j) * 2] = q[j];
            b.dof_state[((size_t)env * NDOF + j) * 2 + 1] = 0.0f;
            act[j] = lact[j] = llact[j] = 0.0f;
        }
#pragma unroll
        for (int k = 0; k < 13; ++k) root[k] = p.base_init_state[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) root[k] += b.env_origins[env * 3 + k];
        if (p.custom_origins) {
            root[0] += p.reset_xy_span * u[10] + p.reset_xy_lo;
            root[1] += p.reset_xy_span * u[11] + p.reset_xy_lo;
        }
#pragma unroll
        for (int k = 0; k < 13; ++k) b.root_states[(size_t)env * 13 + k] = root[k];
        cmd[0] = p.cmd_span[0] * u[12] + p.cmd_lo[0];
        cmd[1] = p.cmd_span[1] * u[13] + p.cmd_lo[1];
        cmd[3] = p.cmd_span[2] * u[14] + p.cmd_lo[2];
        const float keep = (sqrtf(cmd[0] * cmd[0] + cmd[1] * cmd[1]) > 0.2f) ? 1.0f : 0.0f;
        cmd[0] *= keep, cmd[1] *= keep;
        air[0] = air[1] = 0.0f;
        ep_len = 0;
#pragma unroll
        for (int k = 0; k < HB_NUM_REWARDS; ++k) sums[k] = 0.0f;
        grav = quat_rotate_inverse(root + 3, {0.0f, 0.0f, -1.0f});
        eul = euler_xyz_wrapped(root + 3);
    }

    // ---------------- compute_observations, newest frames (hector_env.py:172-254) ----------------
    const bool emit_obs = do_step || (stages & HB_STAGE_OBS);
    if (emit_obs) {
        const float phase = ((float)ep_len * dt) / p.cycle_time;
        const float arg = TWO_PI_F * phase;
        const float s = sinf(arg), c = cosf(arg);
        float st[2] = {(s >= 0.0f) ? 1.0f : 0.0f, (s < 0.0f) ? 1.0f : 0.0f};
        if (fabsf(s) < 0.1f) st[0] = st[1] = 1.0f;
        float fr[PRIV];
        fr[0] = s, fr[1] = c;
        fr[2] = cmd[0] * p.obs_lin_vel, fr[3] = cmd[1] * p.obs_lin_vel, fr[4] = cmd[2] * p.obs_ang_vel;
#pragma unroll
        for (int j = 0; j < NDOF; ++j) {
            fr[5 + j] = (q[j] - p.default_dof_pos[j]) * p.obs_dof_pos;
            fr[5 + NDOF + j] = qd[j] * p.obs_dof_vel;
            fr[5 + 2 * NDOF + j] = act[j];
        }
        constexpr int B0 = 5 + 3 * NDOF;
        const float clip = p.clip_observations;
        float *so = sm, *sp = sm + TILE * OBS;
        {   // actor frame: ... ang_vel, euler, plus scaled noise
            float o[OBS];
#pragma unroll
            for (int k = 0; k < B0; ++k) o[k] = fr[k];
            o[B0] = ang.x * p.obs_ang_vel, o[B0 + 1] = ang.y * p.obs_ang_vel, o[B0 + 2] = ang.z * p.obs_ang_vel;
            o[B0 + 3] = eul.x * p.obs_quat, o[B0 + 4] = eul.y * p.obs_quat, o[B0 + 5] = eul.z * p.obs_quat;
#pragma unroll
            for (int k = 0; k < OBS; ++k) {
                float v = o[k];
                if (with_noise) v = v + (zobs[k] * p.noise_scale_vec[k]) * p.noise_level;
                so[lane * OBS + k] = clampf(v, -clip, clip);
            }
        }
        fr[B0] = lin.x * p.obs_lin_vel, fr[B0 + 1] = lin.y * p.obs_lin_vel, fr[B0 + 2] = lin.z * p.obs_lin_vel;
        fr[B0 + 3] = ang.x * p.obs_ang_vel, fr[B0 + 4] = ang.y * p.obs_ang_vel, fr[B0 + 5] = ang.z * p.obs_ang_vel;
        fr[B0 + 6] = eul.x * p.obs_quat, fr[B0 + 7] = eul.y * p.obs_quat, fr[B0 + 8] = eul.z * p.obs_quat;
#pragma unroll
        for (int f = 0; f < 2; ++f)
#pragma unroll
            for (int k = 0; k < 3; ++k) fr[B0 + 9 + f * 3 + k] = foot_pos[f][k], fr[B0 + 15 + f * 3 + k] = foot_vel[f][k];
        fr[B0 + 21] = root[0], fr[B0 + 22] = root[1], fr[B0 + 23] = root[2];
        fr[B0 + 24] = push_f[0], fr[B0 + 25] = push_f[1];
        fr[B0 + 26] = push_t[0], fr[B0 + 27] = push_t[1], fr[B0 + 28] = push_t[2];
        fr[B0 + 29] = valid ? b.env_frictions[env] : 0.0f;
        fr[B0 + 30] = (valid ? b.body_mass[env] : 0.0f) / 30.0f;
        fr[B0 + 31] = st[0], fr[B0 + 32] = st[1];
        fr[B0 + 33] = (foot_f[0][2] > 5.0f) ? 1.0f : 0.0f, fr[B0 + 34] = (foot_f[1][2] > 5.0f) ? 1.0f : 0.0f;
#pragma unroll
        for (int k = 0; k < PRIV; ++k) sp[lane * PRIV + k] = clampf(fr[k], -clip, clip);
    }
    __syncwarp();
    if (emit_obs) {   // coalesced-by-row store of the newest frames into the last slot of the stacked buffers
        const int ostride = p.frame_stack * OBS, pstride = p.c_frame_stack * PRIV;
        const float *so = sm, *sp = sm + TILE * OBS;
        for (int i = lane; i < nv * OBS; i += 32) {
            const int r = i / OBS, c = i - r * OBS;
            obs_new[(size_t)(env0 + r) * ostride + (ostride - OBS) + c] = so[i];
        }
        for (int i = lane; i < nv * PRIV; i += 32) {
            const int r = i / PRIV, c = i - r * PRIV;
            priv_new[(size_t)(env0 + r) * pstride + (pstride - PRIV) + c] = sp[i];
        }
    }

    // ---------------- state write-back (incl. legged_robot.py:146-150) ----------------
    if (valid) {
#pragma unroll
        for (int j = 0; j < NDOF; ++j) {
            b.last_last_actions[(size_t)env * NDOF + j] = do_step ? lact[j] : llact[j];
            b.last_actions[(size_t)env * NDOF + j] = do_step ? act[j] : lact[j];
            b.last_dof_vel[(size_t)env * NDOF + j] = (do_step || reset) ? qd[j] : ldv[j];
            if (reset) b.actions[(size_t)env * NDOF + j] = 0.0f;
        }
        if (do_step || (stages & HB_STAGE_DERIVE)) {
            b.base_lin_vel[env * 3] = lin.x, b.base_lin_vel[env * 3 + 1] = lin.y, b.base_lin_vel[env * 3 + 2] = lin.z;
            b.base_ang_vel[env * 3] = ang.x, b.base_ang_vel[env * 3 + 1] = ang.y, b.base_ang_vel[env * 3 + 2] = ang.z;
        }
        if (do_step) {
#pragma unroll
            for (int k = 0; k < 6; ++k) b.last_root_vel[(size_t)env * 6 + k] = root[7 + k];
            b.rew_buf[env] = rew;
            b.time_out_buf[env] = time_out ? 1 : 0;
#pragma unroll
            for (int f = 0; f < 2; ++f) {
                b.last_contacts[env * 2 + f] = last_ct[f] ? 1 : 0;
                b.feet_height[env * 2 + f] = fh[f];
                b.last_feet_z[env * 2 + f] = lz[f];
            }
        }
        b.projected_gravity[env * 3] = grav.x, b.projected_gravity[env * 3 + 1] = grav.y, b.projected_gravity[env * 3 + 2] = grav.z;
        b.base_euler_xyz[env * 3] = eul.x, b.base_euler_xyz[env * 3 + 1] = eul.y, b.base_euler_xyz[env * 3 + 2] = eul.z;
#pragma unroll
        for (int k = 0; k < 4; ++k) b.commands[(size_t)env * 4 + k] = cmd[k];
        b.feet_air_time[env * 2] = air[0], b.feet_air_time[env * 2 + 1] = air[1];
        b.episode_length_buf[env] = ep_len;
        b.reset_buf[env] = reset ? 1 : 0;
#pragma unroll
        for (int k = 0; k < HB_NUM_REWARDS; ++k) b.episode_sums[(size_t)k * N + env] = sums[k];
    }

    // ---------------- ordered compaction of the reset ids + episode means ----------------
    // Each warp publishes its ballot and the per-term sums of its reset envs; the last CTA to
    // finish scans the ballots in tile order (ascending env ids, like reset_buf.nonzero()).
    if (ballot) {
#pragma unroll
        for (int k = 0; k < HB_NUM_REWARDS; ++k) {
            float v = my_sums[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) b.scratch_partials[(size_t)blockIdx.x * HB_NUM_REWARDS + k] = v;
        }
    }
    if (lane == 0) {
        b.scratch_ballots[blockIdx.x] = ballot;
        __threadfence();
        const unsigned t = atomicAdd(b.scratch_ticket, 1u);
        s_is_last = (t == gridDim.x - 1);
    }
    __syncwarp();
    if (s_is_last) {
        __threadfence();
        const int tiles = gridDim.x;
        int base = 0;
        double acc[HB_NUM_REWARDS];
#pragma unroll
        for (int k = 0; k < HB_NUM_REWARDS; ++k) acc[k] = 0.0;
        for (int t0 = 0; t0 < tiles; t0 += 32) {
            const int t = t0 + lane;
            const unsigned m = (t < tiles) ? __ldcg(b.scratch_ballots + t) : 0u;
            const int cnt = __popc(m);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            int off = base + incl - cnt;
            unsigned mm = m;
            while (mm) {
                const int bit = __ffs(mm) - 1;
                mm &= mm - 1;
                b.reset_env_ids[off++] = t * TILE + bit;
            }
            if (m) {
#pragma unroll
                for (int k = 0; k < HB_NUM_REWARDS; ++k)
                    acc[k] += (double)__ldcg(b.scratch_partials + (size_t)t * HB_NUM_REWARDS + k);
            }
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
        // extras["episode"] is only refreshed on steps with >= 1 reset (quirk 4): otherwise the
        // previous values are carried into this step's slot.
#pragma unroll
        for (int k = 0; k < HB_NUM_REWARDS; ++k) {
            double v = acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) {
                if (base > 0) b.episode_means[k] = (float)(v / (double)base) / p.max_episode_length_s;
                else if (b.episode_means_prev && b.episode_means_prev != b.episode_means)
                    b.episode_means[k] = b.episode_means_prev[k];
            }
        }
        if (lane == 0) {
            *b.reset_count = base;
            if (host_count) *host_count = base;
            *b.scratch_ticket = 0u;      // re-arm for the next launch
        }
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
__global__ void __launch_bounds__(256)
stack_shift_kernel(const float *__restrict__ prev, float *__restrict__ next, const uint8_t *__restrict__ reset_buf,
                   long long total_vec, int row, int frame, const uint8_t *__restrict__ latch_src,
                   uint8_t *__restrict__ latch_dst, const int32_t *__restrict__ reset_count, int num_envs) {
    const int lane = threadIdx.x & 31;
    const int keep = row - frame;                       // floats of a row that are carried over
    const int rot = frame & 3;                          // source misalignment in floats
    const int fvec = frame >> 2;                        // whole vectors of shift
    const long long warp_base = ((long long)blockIdx.x * blockDim.x + threadIdx.x - lane) * UNROLL;
    // extras["time_outs"] latch (legged_robot.py:208-209 only runs when >= 1 env was reset)
    if (latch_dst && *reset_count > 0) {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < num_envs;
             i += (long long)gridDim.x * blockDim.x)
            latch_dst[i] = latch_src[i];
    }
    const float4 *p4 = reinterpret_cast<const float4 *>(prev);
    float4 *n4 = reinterpret_cast<float4 *>(next);
    float4 v[UNROLL];
    float nx31[UNROLL][3];
    // all loads first (UNROLL independent 16-byte requests in flight per thread)
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const long long i = warp_base + (long long)u * 32 + lane;      // destination vector index
        const long long s = i + fvec;                                   // aligned source vector below the window
        v[u] = (s < total_vec) ? hb::ld_stream4(p4 + s) : make_float4(0.f, 0.f, 0.f, 0.f);
        nx31[u][0] = nx31[u][1] = nx31[u][2] = 0.0f;
        if (lane == 31 && rot != 0 && s + 1 < total_vec) {              // no neighbour lane: fetch the spill-over
            const float *e = prev + (s + 1) * 4;
            nx31[u][0] = __ldg(e);
            if (rot > 1) nx31[u][1] = __ldg(e + 1);
            if (rot > 2) nx31[u][2] = __ldg(e + 2);
        }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const long long i = warp_base + (long long)u * 32 + lane;
        float nx0 = __shfl_down_sync(0xffffffffu, v[u].x, 1);           // first floats of vector s+1
        float nx1 = __shfl_down_sync(0xffffffffu, v[u].y, 1);
        float nx2 = __shfl_down_sync(0xffffffffu, v[u].z, 1);
        if (lane == 31) nx0 = nx31[u][0], nx1 = nx31[u][1], nx2 = nx31[u][2];
        if (i >= total_vec) continue;
        float4 o;
        if (rot == 0) o = v[u];
        else if (rot == 1) o = make_float4(v[u].y, v[u].z, v[u].w, nx0);
        else if (rot == 2) o = make_float4(v[u].z, v[u].w, nx0, nx1);
        else o = make_float4(v[u].w, nx0, nx1, nx2);
        const long long e0 = i * 4;                      // first destination float
        const int r0 = (int)(e0 / row);
        const int c0 = (int)(e0 - (long long)r0 * row);
        if (c0 + 3 < keep) {                             // whole vector inside the carried part of one row
            if (reset_buf[r0]) o = make_float4(0.f, 0.f, 0.f, 0.f);
            hb::st_stream4(n4 + i, o);
        } else {                                         // touches the newest-frame hole or a row boundary
            const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int r = r0, c = c0 + k;
                if (c >= row) c -= row, r += 1;
                if (c < keep) next[e0 + k] = reset_buf[r] ? 0.0f : ov[k];
            }
        }
    }
}

// generic fallback (rows not a multiple of 4 floats in total, or unaligned buffers)
__global__ void __launch_bounds__(256)
stack_shift_scalar_kernel(const float *__restrict__ prev, float *__restrict__ next,
                          const uint8_t *__restrict__ reset_buf, long long total, int row, int frame) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
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
        post_physics_kernel<true><<<tiles, TILE, smem, (cudaStream_t)stream>>>(*p, *buf, *noise, obs_new, priv_new,
                                                                               stages, host_count);
    } else {
        if (!attr_set[0]) {
            HB_CUDA(cudaFuncSetAttribute(post_physics_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            attr_set[0] = true;
        }
        post_physics_kernel<false><<<tiles, TILE, smem, (cudaStream_t)stream>>>(*p, *buf, *noise, obs_new, priv_new,
                                                                                stages, host_count);
    }
    HB_CHECK_LAUNCH("post_physics_kernel");
    return HB_OK;
}

static int launch_stack(const float *prev, float *next, const uint8_t *reset_buf, int n, int row, int frame,
                        const uint8_t *latch_src, uint8_t *latch_dst, const int32_t *reset_count, cudaStream_t st) {
    const long long total = (long long)n * row;
    if ((total % 4) == 0 && hb::aligned16(prev) && hb::aligned16(next)) {
        constexpr int UNROLL = 4;
        const long long total_vec = total / 4;
        const long long threads = (total_vec + UNROLL - 1) / UNROLL;
        const int blocks = (int)((threads + 255) / 256);
        stack_shift_kernel<UNROLL><<<blocks, 256, 0, st>>>(prev, next, reset_buf, total_vec, row, frame, latch_src,
                                                            latch_dst, reset_count, n);
    } else {
        stack_shift_scalar_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(prev, next, reset_buf, total, row, frame);
    }
    return 0;
}

int hb_env_stack_observations(const hb_env_params *p, const hb_env_buffers *buf, const float *obs_prev,
                              const float *priv_prev, float *obs_new, float *priv_new, void *stream) {
    if (int rc = check_params(p, buf, "hb_env_stack_observations")) return rc;
    HB_REQUIRE(obs_prev && priv_prev && obs_new && priv_new && buf->reset_buf, "hb_env_stack_observations: null buffer");
    HB_REQUIRE(obs_prev != obs_new && priv_prev != priv_new, "hb_env_stack_observations: prev and new must not alias");
    cudaStream_t st = (cudaStream_t)stream;
    launch_stack(obs_prev, obs_new, buf->reset_buf, p->num_envs, p->frame_stack * p->num_single_obs, p->num_single_obs,
                 buf->time_out_buf, buf->time_outs_latched, buf->reset_count, st);
    HB_CHECK_LAUNCH("stack_shift_kernel(obs)");
    launch_stack(priv_prev, priv_new, buf->reset_buf, p->num_envs, p->c_frame_stack * p->num_single_priv,
                 p->num_single_priv, nullptr, nullptr, nullptr, st);
    HB_CHECK_LAUNCH("stack_shift_kernel(priv)");
    return HB_OK;
}

}  // extern "C"
