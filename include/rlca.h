/*
 * rlca.h — C ABI of the B200-native collision-avoidance hot path (librlca.so).
 *
 * Drop-in boundary for the path BASELINE.json's north_star names.  The
 * reference has no FFI of its own (its seam is duck-typed Python over ROS
 * topics); each entry point below states the reference
 * interface it replaces (file:line under /root/reference).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; no torch / C++ types.
 *   - Every function returns int: 0 = RLCA_OK, otherwise an rlca_status;
 *     the message is available from rlca_last_error() (thread-local).
 *   - Pointers named *_dev are DEVICE pointers owned by the caller (PyTorch
 *     allocates them); the library never frees them.  `stream` is a
 *     cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *     asynchronous on that stream unless the name ends in _host.
 *   - A handle is bound to the CUDA device current at creation, owns only the
 *     uploaded static map, scenario tables and the beam table, and is not
 *     thread-safe.  Different handles are independent.
 *   - There is no CPU fallback: without a CUDA device every call fails loudly.
 */
#ifndef RLCA_H
#define RLCA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum rlca_status {
    RLCA_OK = 0,
    RLCA_ERR_INVALID = 1,     /* bad argument / config */
    RLCA_ERR_CUDA = 2,        /* CUDA runtime error (sticky errors surface here) */
    RLCA_ERR_UNSUPPORTED = 3, /* e.g. map too large for the shared-memory owner grid */
    RLCA_ERR_NO_DEVICE = 4
} rlca_status;

#define RLCA_MAX_ROBOTS_PER_WORLD 64

/* Scenario + geometry.  Field meanings and reference sources:
 *   robots_per_world  24/44/50 agents per world   worlds/stage1.world:107-130, stage2.world:113-165, circle.world:106-155
 *   beams / raw_beams beam_num / sensor samples   stage_world1.py:17,126-139 ; worlds/stage1.world:14
 *   resolution        cell size                   worlds/stage1.world:3
 *   dt                0.1 s tick                  default interval_sim; test/hztest.xml:14,18
 *   range_max, fov    6.0 m, pi                   worlds/stage1.world:12-13
 *   half_len/half_wid 0.22, 0.19                  worlds/stage1.world:83
 *   goal_radius ... w_penalty                     stage_world1.py:34,183-204 ; circle_world.py:195
 *   v_min..w_max      action bound                ppo_stage1.py:170
 *   timeout           150/200/10000               stage_world1.py:206, stage_world2.py:203, circle_world.py:198
 *   pre_distance_zero quirk                       stage_world2.py:170, circle_world.py:166
 *   scenario          0 = stage1 random spawn/goal (stage_world1.py:251-274)
 *                     1 = stage2 tables + random region for flagged rows (stage_world2.py:164-171,210-221,250-287)
 *                     2 = circle tables (circle_world.py:164-167,205-208)
 */
typedef struct rlca_env_config {
    int32_t robots_per_world;
    int32_t num_worlds;          /* worlds on THIS device (shard) */
    int32_t beams;
    int32_t raw_beams;
    int32_t grid_w, grid_h;      /* static map cells; grid_w is also the row pitch */
    int32_t origin_cx, origin_cy;/* cell index of world (0,0): cell = floor(x*ppm) + origin */
    float resolution;
    float ppm;
    float dt;
    float inv_dt;
    float range_max;
    float range_cells;           /* ppm * range_max */
    float fov;
    float half_len, half_wid;
    float goal_radius;
    float reward_arrive;
    float reward_collision;
    float progress_gain;
    float w_threshold;
    float w_penalty;
    float v_min, v_max, w_min, w_max;
    int32_t timeout;
    int32_t pre_distance_zero;
    int32_t scenario;
    int32_t auto_reset;          /* done agents are re-spawned inside the step kernel */
    int32_t max_reject;          /* cap on rejection-sampling tries */
    int32_t world_offset;        /* global index of this shard's first world (RNG keys are global) */
    uint64_t seed;
} rlca_env_config;

/* Per-agent simulator state, N = robots_per_world * num_worlds rows of 4 x 32 bit:
 *   pose  x, y, theta, distance-to-goal after the last tick (= pre_distance next tick)
 *   goal  goal_x, goal_y, last commanded v, last commanded w (odom twist, stageros.cpp:547-550)
 *   acc   episode reward, last reward, init_x, init_y
 *   meta  step counter t, episode index, stall flag (is_crashed, stageros.cpp:560-564), terminal latch */
typedef struct rlca_env_state {
    float *pose_dev;
    float *goal_dev;
    float *acc_dev;
    int32_t *meta_dev;
} rlca_env_state;

/* Inputs / outputs of one tick.  obs may point into a rollout buffer slice.
 *   action (N,2) raw policy action (clipped inside)          model/ppo.py:73-75, stage_world1.py:226-234
 *   live   (N) u8 or NULL; 0 = agent idles on its last command  ppo_stage2.py:72-84
 *   obs    (N,beams) scan/6 - 0.5                             stage_world1.py:122-140
 *   reward (N)  flags (N,4) u8 = done, crashed, result{0,1 Reach Goal,2 Crashed,3 Time out}, was_reset
 *   gs     (N,4) local goal x,y + speed v,w                   stage_world1.py:143-144,155-160
 *   eplog  (N,8) written for agents whose episode ended: goal_x, goal_y, ep_reward, steps,
 *          init_x, init_y, result, episode                    ppo_stage1.py:127-131 */
typedef struct rlca_step_io {
    const float *action_dev;
    const uint8_t *live_dev;
    float *obs_dev;
    float *reward_dev;
    uint8_t *flags_dev;
    float *gs_dev;
    float *eplog_dev;
} rlca_step_io;

typedef struct rlca_env rlca_env;

/* Replaces StageNode construction + world->Load (stageros.cpp:311-355): creates the
 * device-side world for a batch of identical worlds. */
int rlca_env_create(const rlca_env_config *cfg, rlca_env **out);
int rlca_env_destroy(rlca_env *env);

/* Upload the static occupancy grid (HOST pointer, grid_h*grid_w bytes, 0 = free,
 * non-zero = obstacle).  Replaces libstage's bitmap/polygon block rasterisation at load
 * (worlds/stage1.world:43-49, stage2.world:169-297). */
int rlca_env_set_map(rlca_env *env, const uint8_t *cells_host, int32_t grid_w, int32_t grid_h);

/* Scenario tables (HOST pointers, robots_per_world rows of 4 floats):
 *   init_tab  x, y, theta, random_flag   (world-file agent poses / model/utils.py:6-25,41-53)
 *   goal_tab  gx, gy, random_flag, 0     (model/utils.py:27-38,55-63) */
int rlca_env_set_tables(rlca_env *env, const float *init_tab_host, const float *goal_tab_host);

/* reset_world (stage_world1.py:162-169 -> cb_reset_srv stageros.cpp:260-269) when
 * clear_world != 0, then reset_pose + generate_goal_point (stage_world1.py:171-177,213-223)
 * for agents with mask != 0 (mask_dev NULL = all agents).  In place. */
int rlca_env_reset(rlca_env *env, const rlca_env_state *state, const uint8_t *mask_dev,
                   int32_t clear_world, void *stream);

/* Scan / local goal / speed from the current poses without ticking
 * (get_laser_observation, get_local_goal, get_self_speed right after a reset:
 * ppo_stage1.py:59-63).  io->obs_dev and io->gs_dev are written. */
int rlca_env_observe(rlca_env *env, const rlca_env_state *state, const rlca_step_io *io, void *stream);

/* ONE fused tick over the whole agent batch: control_vel (stage_world1.py:226-234) ->
 * World::UpdateAll (stageros.cpp:448: integrate, collide, stall) -> WorldCallback
 * (stageros.cpp:451-611: GT velocity, is_crashed) -> get_reward_and_terminate
 * (stage_world1.py:180-211) -> optional re-spawn -> lidar raytrace from the final pose
 * (stageros.cpp:479-516) -> get_laser_observation / get_local_goal / get_self_speed.
 * Reads state_in, writes state_out (they may alias only when the launch uses one CTA
 * per world; pass distinct buffers and swap them each tick otherwise). */
int rlca_env_step(rlca_env *env, const rlca_env_state *state_in, const rlca_env_state *state_out,
                  const rlca_step_io *io, void *stream);

/* Same tick driven from HOST buffers (the reference-facing call: actions arrive from
 * the host, observations/rewards/flags return to it).  H2D of action_host, the tick,
 * D2H of obs/reward/flags/gs, then a stream synchronize.  Any *_host may be NULL to
 * skip that copy.  io holds the device staging buffers. */
int rlca_env_step_host(rlca_env *env, const rlca_env_state *state_in, const rlca_env_state *state_out,
                       const rlca_step_io *io, const float *action_host, float *obs_host,
                       float *reward_host, uint8_t *flags_host, float *gs_host, void *stream);

/* Stand-alone lidar raycast (World::Raytrace via ModelRanger, stageros.cpp:479-516):
 * pose_dev (N,4) x,y,theta,_ -> ranges_dev (N,beams) in metres (normalise = 0) or
 * scan/6-0.5 (normalise = 1).  Other robots' footprints are seen, own is excluded. */
int rlca_raycast(rlca_env *env, const float *pose_dev, float *ranges_dev, int32_t normalise, void *stream);

/* Launch shape knob: CTAs per world (>= 1).  0 = library default (auto). */
int rlca_env_set_ctas_per_world(rlca_env *env, int32_t ctas_per_world);
/* Number of kernels the library launched on behalf of this handle so far. */
int64_t rlca_env_launch_count(const rlca_env *env);

/* sizeof(rlca_env_config) as compiled, so bindings can verify their struct layout. */
int rlca_sizeof_env_config(void);

const char *rlca_last_error(void);
const char *rlca_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RLCA_H */
