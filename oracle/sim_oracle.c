/*
 * sim_oracle.c — CPU restatement (the ORACLE) of the reference's simulator hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (rl_collision_avoidance_b200/) never links, imports or calls anything in oracle/.
 *
 * PARITY STATUS: "parity unpinned" for the simulator half.  The arithmetic the
 * reference relies on lives in libstage (github.com/rtv/Stage, un-vendored, no
 * pinned version; ROS Kinetic ships 4.1.1), which is absent from /root/reference
 * and cannot be built here.  The reference's own tests for this path
 * (stage_ros-add_pose_and_crash/test/cmdpose_tests.py:87-203, hztest.xml:8-42)
 * pin only qualitative facts — forward motion moves along the heading only, yaw
 * commands change only yaw, teleport sets (x,y,yaw) exactly, tick = 0.1 s — and
 * those ARE asserted against this oracle in tests/test_oracle_pins.py.
 * Everything else restates the published Stage algorithm as
 * called from the reference's call sites, cited per function below.
 *
 * Numerics contract (shared *specification*, independently implemented on the
 * GPU): all arithmetic is IEEE fp32 with the fused-multiply-adds written
 * explicitly (fmaf) and no other contraction (build with -ffp-contract=off),
 * sin/cos of the heading come from orc_sincosf() below (Cody-Waite + cephes
 * polynomials, written out), beam directions come from a host table in double
 * rounded to float and are rotated by the heading.  That makes poses, ranges,
 * rewards and flags reproducible bit-for-bit by any implementation of the spec.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_ROBOTS 64
#define ORC_MAX_OUTLINE 1024

/* Mirrors rlca_env_config field-for-field in meaning (include/rlca.h) but is
 * declared independently: the oracle shares no header with the product. */
typedef struct {
    int32_t robots_per_world; /* 24 / 44 / 50: worlds/stage1.world:107-130 etc. */
    int32_t num_worlds;
    int32_t beams;            /* rays cast per robot (beam_num) */
    int32_t raw_beams;        /* sensor samples (512: worlds/stage1.world:14) */
    int32_t grid_w, grid_h;   /* cells */
    int32_t origin_cx, origin_cy; /* cell index of world (0,0): cell = floor(x*ppm)+origin */
    float resolution;         /* cell size, worlds/stage1.world:3 */
    float ppm;                /* 1/resolution as float */
    float dt;                 /* 0.1 s: default interval_sim, hztest.xml:14,18 */
    float inv_dt;             /* 1.0f/dt */
    float range_max;          /* 6.0: worlds/stage1.world:13 */
    float range_cells;        /* ppm*range_max as float */
    float fov;                /* pi (180 deg): worlds/stage1.world:12 */
    float half_len, half_wid; /* 0.22, 0.19: size [0.44 0.38], worlds/stage1.world:83 */
    float goal_radius;        /* 0.5: stage_world1.py:34 */
    float reward_arrive;      /* 15: stage_world1.py:187 */
    float reward_collision;   /* -15: stage_world1.py:195 */
    float progress_gain;      /* 2.5: stage_world1.py:183 */
    float w_threshold;        /* 1.05 (0.7 circle): stage_world1.py:203, circle_world.py:195 */
    float w_penalty;          /* -0.1 */
    float v_min, v_max, w_min, w_max; /* action bound, ppo_stage1.py:170 */
    int32_t timeout;          /* t > 150 / 200 / 10000 */
    int32_t pre_distance_zero;/* stage_world2.py:170, circle_world.py:166 quirk */
    int32_t scenario;         /* 0 stage1 random; 1 stage2 table+random; 2 circle table */
    int32_t auto_reset;
    int32_t max_reject;       /* cap on rejection-sampling tries */
    int32_t world_offset;     /* global index of this shard's first world (multi-GPU) */
    uint64_t seed;
} orc_config;

/* ------------------------------------------------------------------------- */
/* deterministic fp32 sin/cos (spec: Cody-Waite 2-term reduction by pi/2,    */
/* cephes sinf/cosf minimax polynomials on [-pi/4,pi/4], explicit fmaf).      */
void orc_sincosf(float x, float *s, float *c)
{
    const float two_over_pi = 0.636619772367581343f;
    const float pio2_hi = 1.57079625129699707031f;   /* 0x3FC90FDA */
    const float pio2_lo = 7.54978941586159635335e-08f;
    float q = rintf(x * two_over_pi);
    float r = fmaf(q, -pio2_hi, x);
    r = fmaf(q, -pio2_lo, r);
    float r2 = r * r;
    /* sin(r) = r + r^3 * (S1 + r2*(S2 + r2*S3)) */
    float ps = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(r2, ps, -1.6666654611e-1f);
    float sr = fmaf(r * r2, ps, r);
    /* cos(r) = 1 - r2/2 + r2^2 * (C1 + r2*(C2 + r2*C3)) */
    float pc = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(r2, pc, 4.166664568298827e-2f);
    float cr = fmaf(r2 * r2, pc, fmaf(r2, -0.5f, 1.0f));
    int qi = ((int)q) & 3;
    float ss, cc;
    switch (qi) {
    case 0: ss = sr; cc = cr; break;
    case 1: ss = cr; cc = -sr; break;
    case 2: ss = -sr; cc = -cr; break;
    default: ss = -cr; cc = sr; break;
    }
    *s = ss; *c = cc;
}

/* Stg::normalize() restated for one wrap (|increment| <= pi): result in (-pi, pi]. */
static float orc_normalize(float a)
{
    const float pi_f = 3.14159274101257324219f;
    const float two_pi_f = 6.28318548202514648438f;
    if (a > pi_f) a -= two_pi_f;
    else if (a <= -pi_f) a += two_pi_f;
    return a;
}

/* ------------------------------------------------------------------------- */
/* Philox4x32-10 counter RNG (Salmon et al. 2011).  Replaces the reference's  */
/* unseeded np.random.uniform (stage_world1.py:252-272) with a keyed stream:  */
/* key = seed, counter = (global agent id, episode, draw index, purpose).     */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int i = 0; i < 10; ++i) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

/* 4 uniforms in [0,1) with 24 bits each */
static void orc_rand4(uint64_t seed, uint32_t agent, uint32_t episode, uint32_t draw,
                      uint32_t purpose, float u[4])
{
    uint32_t c[4] = { agent, episode, draw, purpose };
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    for (int i = 0; i < 4; ++i) u[i] = (float)(c[i] >> 8) * 5.9604644775390625e-08f;
}

static float orc_uniform(float u, float lo, float hi) { return fmaf(u, hi - lo, lo); }

/* ------------------------------------------------------------------------- */
/* Cohen integer line (Graphics Gems IV) as used by Stage for both block      */
/* rasterisation (World::ForEachCellInLine) and World::Raytrace — SURVEY.md   */
/* App. A.3 / A.7.  Marks n = ax+ay cells starting at (x0,y0), end excluded.  */
static int line_cells(int x0, int y0, int x1, int y1, int *out_x, int *out_y, int cap)
{
    int dx = x1 - x0, dy = y1 - y0;
    int sx = (dx > 0) - (dx < 0), sy = (dy > 0) - (dy < 0);
    int ax = abs(dx), ay = abs(dy);
    int bx = 2 * ax, by = 2 * ay;
    int exy = ay - ax;
    int n = ax + ay;
    int gx = x0, gy = y0, cnt = 0;
    while (n > 0) {
        if (cnt < cap) { out_x[cnt] = gx; out_y[cnt] = gy; }
        ++cnt;
        if (exy < 0) { gx += sx; exy += by; }
        else { gy += sy; exy -= bx; }
        --n;
    }
    return cnt;
}

/* Robot footprint outline -> global cells.  Unit-square block scaled to
 * size [0.44 0.38], centred (origin [0 0 0 0]), rotated by the heading
 * (worlds/stage1.world:83,87); only polygon EDGES are rasterised (App. A.3). */
static int robot_outline(const orc_config *cfg, float x, float y, float th, int *ox, int *oy)
{
    float s, c;
    orc_sincosf(th, &s, &c);
    const float hx[4] = { -cfg->half_len, cfg->half_len, cfg->half_len, -cfg->half_len };
    const float hy[4] = { -cfg->half_wid, -cfg->half_wid, cfg->half_wid, cfg->half_wid };
    int cx[4], cy[4];
    for (int k = 0; k < 4; ++k) {
        float px = fmaf(hx[k], c, fmaf(-hy[k], s, x));
        float py = fmaf(hx[k], s, fmaf(hy[k], c, y));
        cx[k] = (int)floorf(px * cfg->ppm);
        cy[k] = (int)floorf(py * cfg->ppm);
    }
    int cnt = 0;
    for (int k = 0; k < 4; ++k) {
        int k1 = (k + 1) & 3;
        cnt += line_cells(cx[k], cy[k], cx[k1], cy[k1], ox + cnt, oy + cnt, ORC_MAX_OUTLINE - cnt);
        if (cnt > ORC_MAX_OUTLINE) cnt = ORC_MAX_OUTLINE;
    }
    return cnt;
}

/* Owner grid: 0 empty, 1..R = exactly one robot (id+1), 255 = two or more
 * robots, 254 = static obstacle (wall / polygon obstacle).  A cell blocks
 * robot i (ray or footprint) iff value != 0 && value != i+1: that is Stage's
 * "block of an unrelated model" predicate (App. A.3, A.6) on a cell grid. */
#define CELL_STATIC 254
#define CELL_MULTI 255

static void grid_mark(uint8_t *g, const orc_config *cfg, int gx, int gy, int robot)
{
    int cx = gx + cfg->origin_cx, cy = gy + cfg->origin_cy;
    if (cx < 0 || cy < 0 || cx >= cfg->grid_w || cy >= cfg->grid_h) return;
    uint8_t *p = &g[cy * cfg->grid_w + cx];
    uint8_t me = (uint8_t)(robot + 1);
    if (*p == 0) *p = me;
    else if (*p != me && *p != CELL_STATIC) *p = CELL_MULTI;
}

static int grid_blocks(const uint8_t *g, const orc_config *cfg, int gx, int gy, int robot)
{
    int cx = gx + cfg->origin_cx, cy = gy + cfg->origin_cy;
    if (cx < 0 || cy < 0 || cx >= cfg->grid_w || cy >= cfg->grid_h) return 0; /* outside = empty */
    uint8_t v = g[cy * cfg->grid_w + cx];
    return v != 0 && v != (uint8_t)(robot + 1);
}

/* World::Raytrace restated (consumer stageros.cpp:479-516):
 * pure Cohen walk, no empty-region jumps (documented deviation), self-excluding. */
static float cast_ray(const uint8_t *g, const orc_config *cfg, int robot,
                      int gx0, int gy0, float ca, float sa)
{
    float dx = cfg->range_cells * ca;
    float dy = cfg->range_cells * sa;
    int idx = (int)dx, idy = (int)dy;          /* truncation toward zero */
    int sx = (dx > 0.0f) - (dx < 0.0f), sy = (dy > 0.0f) - (dy < 0.0f);
    int ax = abs(idx), ay = abs(idy);
    int bx = 2 * ax, by = 2 * ay;
    int exy = ay - ax;
    int n = ax + ay;
    int gx = gx0, gy = gy0;
    while (n > 0) {
        if (grid_blocks(g, cfg, gx, gy, robot)) {
            if (ax > ay) return fabsf((float)(gx - gx0) / ca) * cfg->resolution;
            else         return fabsf((float)(gy - gy0) / sa) * cfg->resolution;
        }
        if (exy < 0) { gx += sx; exy += by; }
        else { gy += sy; exy -= bx; }
        --n;
    }
    return cfg->range_max;
}

/* Beam table: bearing_i = -fov/2 + i*fov/(raw-1) (stageros.cpp:495-497,
 * App. A.6), sub-sampled with get_laser_observation's symmetric index map
 * (stage_world1.py:126-139). cos/sin evaluated in double, rounded to float. */
void orc_beam_table(const orc_config *cfg, float *cosb, float *sinb, int32_t *raw_index)
{
    int raw = cfg->raw_beams, nb = cfg->beams;
    double step = (double)raw / (double)nb;
    int half = nb / 2;
    double index = 0.0;
    for (int i = 0; i < half; ++i) { raw_index[i] = (int)index; index += step; }
    index = raw - 1.0;
    for (int i = 0; i < half; ++i) { raw_index[nb - 1 - i] = (int)index; index -= step; }
    for (int i = 0; i < nb; ++i) {
        double b = -0.5 * (double)cfg->fov + (double)raw_index[i] * ((double)cfg->fov / (double)(raw - 1));
        cosb[i] = (float)cos(b);
        sinb[i] = (float)sin(b);
    }
}

static void scan_robot(const uint8_t *g, const orc_config *cfg, int robot,
                       float x, float y, float th,
                       const float *cosb, const float *sinb, float *out, int normalise)
{
    float st, ct;
    orc_sincosf(th, &st, &ct);
    int gx0 = (int)floorf(x * cfg->ppm);
    int gy0 = (int)floorf(y * cfg->ppm);
    for (int b = 0; b < cfg->beams; ++b) {
        float ca = fmaf(ct, cosb[b], -(st * sinb[b]));
        float sa = fmaf(st, cosb[b], ct * sinb[b]);
        float r = cast_ray(g, cfg, robot, gx0, gy0, ca, sa);
        /* get_laser_observation: scan/6.0 - 0.5 (stage_world1.py:140) */
        out[b] = normalise ? fmaf(r, 1.0f / 6.0f, -0.5f) : r;
    }
}

static void build_grid(uint8_t *g, const uint8_t *static_cells, const orc_config *cfg,
                       const float *px, const float *py, const float *pth)
{
    int ox[ORC_MAX_OUTLINE], oy[ORC_MAX_OUTLINE];
    memcpy(g, static_cells, (size_t)cfg->grid_w * cfg->grid_h);
    for (int r = 0; r < cfg->robots_per_world; ++r) {
        int n = robot_outline(cfg, px[r], py[r], pth[r], ox, oy);
        for (int k = 0; k < n; ++k) grid_mark(g, cfg, ox[k], oy[k], r);
    }
}

/* Stand-alone raycast of every robot (static map + other robots' footprints).
 * pose: (N,4) floats x,y,theta,unused. ranges: (N,beams) raw metres. */
void orc_raycast(const orc_config *cfg, const uint8_t *static_cells, const float *pose,
                 float *ranges, int normalise)
{
    int R = cfg->robots_per_world;
    float *cosb = (float *)malloc(sizeof(float) * cfg->beams);
    float *sinb = (float *)malloc(sizeof(float) * cfg->beams);
    int32_t *ri = (int32_t *)malloc(sizeof(int32_t) * cfg->beams);
    orc_beam_table(cfg, cosb, sinb, ri);
#pragma omp parallel
    {
        uint8_t *g = (uint8_t *)malloc((size_t)cfg->grid_w * cfg->grid_h);
        float px[ORC_MAX_ROBOTS], py[ORC_MAX_ROBOTS], pth[ORC_MAX_ROBOTS];
#pragma omp for schedule(dynamic, 1)
        for (int w = 0; w < cfg->num_worlds; ++w) {
            for (int r = 0; r < R; ++r) {
                const float *p = pose + 4 * ((size_t)w * R + r);
                px[r] = p[0]; py[r] = p[1]; pth[r] = p[2];
            }
            build_grid(g, static_cells, cfg, px, py, pth);
            for (int r = 0; r < R; ++r)
                scan_robot(g, cfg, r, px[r], py[r], pth[r], cosb, sinb,
                           ranges + ((size_t)w * R + r) * cfg->beams, normalise);
        }
        free(g);
    }
    free(cosb); free(sinb); free(ri);
}

/* ------------------------------------------------------------------------- */
/* Episode reset: reset_pose / generate_random_pose / generate_goal_point /   */
/* generate_random_goal (stage_world1.py:171-177,213-223,251-274;             */
/* stage_world2.py:164-171,210-221,250-287; circle_world.py:164-167,205-208). */
/* tables: init (R,4) x,y,theta,random_flag ; goal (R,4) gx,gy,random_flag,_  */
static void stage2_random_xy(const orc_config *cfg, uint32_t agent, uint32_t episode,
                             uint32_t purpose, float refx, float refy, float *ox, float *oy,
                             float *otheta)
{
    float u[4];
    float x = 0.f, y = 0.f;
    for (int k = 0; k < cfg->max_reject; ++k) {
        orc_rand4(cfg->seed, agent, episode, (uint32_t)k, purpose, u);
        x = orc_uniform(u[0], 9.0f, 19.0f);
        y = u[1];
        if (y <= 0.4f) y = -fmaf(y, 10.0f, 1.0f);
        else y = -fmaf(y, 10.0f, 9.0f);
        float ddx = x - refx, ddy = y - refy;
        float dis = sqrtf(fmaf(ddx, ddx, ddy * ddy));
        if (!(dis < 7.0f)) break;
    }
    orc_rand4(cfg->seed, agent, episode, 0xFFFFu, purpose, u);
    *ox = x; *oy = y; *otheta = orc_uniform(u[0], 0.0f, 6.28318548202514648438f);
}

/* state arrays are (N,4): pose = x,y,theta,dist ; goal = gx,gy,v_cmd,w_cmd ;
 * acc = ep_reward,last_reward,init_x,init_y ; meta(int32) = t,episode,stall,terminal */
/* goal_only: generate_goal_point alone (stage_world1.py:171-177) - a new goal for the CURRENT pose with the draws of
 * the current episode, pre_distance and init_pose refreshed, counters untouched. */
static void reset_agent(const orc_config *cfg, const float *init_tab, const float *goal_tab,
                        uint32_t agent_gid, int r, float *pose, float *goal, float *acc,
                        int32_t *meta, int goal_only)
{
    uint32_t episode = (uint32_t)(meta[1] + (goal_only ? 0 : 1));
    meta[1] = (int32_t)episode;
    float u[4];
    float x, y, th;
    int random_pose = (cfg->scenario == 0) || (cfg->scenario == 1 && init_tab[4 * r + 3] != 0.0f);
    if (goal_only) {
        x = pose[0]; y = pose[1]; th = pose[2];
    } else if (cfg->scenario == 0) {
        x = y = 0.f;
        for (int k = 0; k < cfg->max_reject; ++k) {
            orc_rand4(cfg->seed, agent_gid, episode, (uint32_t)k, 1u, u);
            x = orc_uniform(u[0], -9.0f, 9.0f);
            y = orc_uniform(u[1], -9.0f, 9.0f);
            float dis = sqrtf(fmaf(x, x, y * y));
            if (!(dis > 9.0f)) break;
        }
        orc_rand4(cfg->seed, agent_gid, episode, 0xFFFFu, 1u, u);
        th = orc_uniform(u[0], 0.0f, 6.28318548202514648438f);
    } else if (random_pose) {
        stage2_random_xy(cfg, agent_gid, episode, 1u, pose[0], pose[1], &x, &y, &th);
    } else {
        x = init_tab[4 * r + 0]; y = init_tab[4 * r + 1]; th = init_tab[4 * r + 2];
    }
    th = orc_normalize(th);  /* yaw read back through a quaternion is in (-pi,pi] (stage_world1.py:90) */
    /* teleport: SetPose, no collision test, stall untouched (stageros.cpp:282-296) */
    pose[0] = x; pose[1] = y; pose[2] = th;
    float gx, gy;
    int random_goal = (cfg->scenario == 0) || (cfg->scenario == 1 && goal_tab[4 * r + 2] != 0.0f);
    if (cfg->scenario == 0) {
        gx = gy = 0.f;
        for (int k = 0; k < cfg->max_reject; ++k) {
            orc_rand4(cfg->seed, agent_gid, episode, (uint32_t)k, 2u, u);
            gx = orc_uniform(u[0], -9.0f, 9.0f);
            gy = orc_uniform(u[1], -9.0f, 9.0f);
            float dis_origin = sqrtf(fmaf(gx, gx, gy * gy));
            float ddx = gx - x, ddy = gy - y;
            float dis_goal = sqrtf(fmaf(ddx, ddx, ddy * ddy));
            if (!(dis_origin > 9.0f || dis_goal > 10.0f || dis_goal < 8.0f)) break;
        }
    } else if (random_goal) {
        float dummy;
        stage2_random_xy(cfg, agent_gid, episode, 2u, x, y, &gx, &gy, &dummy);
    } else {
        gx = goal_tab[4 * r + 0]; gy = goal_tab[4 * r + 1];
    }
    goal[0] = gx; goal[1] = gy;  /* goal[2..3] (last commanded v,w) persist: GetVelocity() */
    float ddx = gx - x, ddy = gy - y;
    float d0 = sqrtf(fmaf(ddx, ddx, ddy * ddy));
    pose[3] = cfg->pre_distance_zero ? 0.0f : d0;
    acc[2] = x; acc[3] = y;  /* init_pose for the 'Distance' log column (ppo_stage1.py:127) */
    if (goal_only) return;
    acc[0] = 0.0f;           /* ep_reward */
    meta[0] = 1;             /* step = 1 (ppo_stage1.py:57) */
    meta[3] = 0;             /* terminal latch cleared */
}

/* reset_world (stage_world1.py:162-169 -> cb_reset_srv stageros.cpp:260-269):
 * world-file poses restored, stall cleared, speeds zeroed.  Then optional per-agent
 * reset_pose + generate_goal_point for agents with mask != 0 (mask NULL = all).
 * clear_world == 2: generate_goal_point only (new goal for the current pose). */
void orc_reset(const orc_config *cfg, const float *init_tab, const float *goal_tab,
               const uint8_t *mask, int clear_world,
               float *pose, float *goal, float *acc, int32_t *meta)
{
    int R = cfg->robots_per_world;
    int N = R * cfg->num_worlds;
    for (int i = 0; i < N; ++i) {
        int r = i % R;
        uint32_t gid = (uint32_t)(cfg->world_offset * R + i);
        if (clear_world == 1) {
            pose[4 * i + 0] = init_tab[4 * r + 0];
            pose[4 * i + 1] = init_tab[4 * r + 1];
            pose[4 * i + 2] = orc_normalize(init_tab[4 * r + 2]);
            pose[4 * i + 3] = 0.0f;
            goal[4 * i + 0] = goal[4 * i + 1] = goal[4 * i + 2] = goal[4 * i + 3] = 0.0f;
            acc[4 * i + 0] = acc[4 * i + 1] = 0.0f;
            acc[4 * i + 2] = pose[4 * i + 0]; acc[4 * i + 3] = pose[4 * i + 1];
            meta[4 * i + 0] = 1; meta[4 * i + 1] = 0; meta[4 * i + 2] = 0; meta[4 * i + 3] = 0;
        }
        if (mask == NULL || mask[i])
            reset_agent(cfg, init_tab, goal_tab, gid, r, pose + 4 * i, goal + 4 * i, acc + 4 * i,
                        meta + 4 * i, clear_world == 2);
    }
}

/* ------------------------------------------------------------------------- */
/* One lock-step tick for every world (SURVEY.md App. A.1 determinised order): */
/* clip action -> integrate -> collide/revert/stall -> GT velocity ->          */
/* reward/done -> (auto-reset) -> raycast from the final pose -> obs.          */
/* live: NULL or (N) u8; a non-live agent holds its last command, re-emits its */
/* last reward with done=1 (ppo_stage2.py:72-84 liveflag semantics).           */
/* outputs: obs (N,beams) normalised; reward (N); flags (N,4) u8 =             */
/* done,crashed,result,was_reset ; gs (N,4) = local_goal_x,y, speed v,w ;      */
/* eplog (N,8) written for done agents: gx,gy,ep_reward,steps,ix,iy,result,ep  */
void orc_step(const orc_config *cfg, const uint8_t *static_cells,
              const float *init_tab, const float *goal_tab,
              const float *action, const uint8_t *live,
              float *pose, float *goal, float *acc, int32_t *meta,
              float *obs, float *reward, uint8_t *flags, float *gs, float *eplog)
{
    int R = cfg->robots_per_world;
    float *cosb = (float *)malloc(sizeof(float) * cfg->beams);
    float *sinb = (float *)malloc(sizeof(float) * cfg->beams);
    int32_t *ri = (int32_t *)malloc(sizeof(int32_t) * cfg->beams);
    orc_beam_table(cfg, cosb, sinb, ri);
#pragma omp parallel
    {
        uint8_t *g = (uint8_t *)malloc((size_t)cfg->grid_w * cfg->grid_h);
        int ox[ORC_MAX_OUTLINE], oy[ORC_MAX_OUTLINE];
        float nx[ORC_MAX_ROBOTS], ny[ORC_MAX_ROBOTS], nth[ORC_MAX_ROBOTS];
        int moving[ORC_MAX_ROBOTS], hit[ORC_MAX_ROBOTS];
#pragma omp for schedule(dynamic, 1)
        for (int w = 0; w < cfg->num_worlds; ++w) {
            size_t base = (size_t)w * R;
            int latch_in[ORC_MAX_ROBOTS];
            for (int r = 0; r < R; ++r) latch_in[r] = meta[4 * (base + r) + 3];
            /* a1/a3: command + explicit-Euler diff-drive with the OLD heading (App. A.2) */
            for (int r = 0; r < R; ++r) {
                size_t i = base + r;
                float v, om;
                if ((live != NULL && !live[i]) || (cfg->auto_reset == 2 && latch_in[r])) { v = goal[4 * i + 2]; om = goal[4 * i + 3]; }
                else {
                    v = action[2 * i + 0]; om = action[2 * i + 1];
                    if (!(fabsf(v) <= 3.0e38f)) v = 0.0f;   /* NaN/Inf guard */
                    if (!(fabsf(om) <= 3.0e38f)) om = 0.0f;
                    v = fminf(fmaxf(v, cfg->v_min), cfg->v_max);   /* np.clip, model/ppo.py:75 */
                    om = fminf(fmaxf(om, cfg->w_min), cfg->w_max);
                    goal[4 * i + 2] = v; goal[4 * i + 3] = om;      /* SetSpeed, stageros.cpp:276 */
                }
                float x = pose[4 * i + 0], y = pose[4 * i + 1], th = pose[4 * i + 2];
                moving[r] = (v != 0.0f) || (om != 0.0f);
                if (moving[r]) {
                    float s, c;
                    orc_sincosf(th, &s, &c);
                    float d = v * cfg->dt;
                    nx[r] = fmaf(d, c, x);
                    ny[r] = fmaf(d, s, y);
                    nth[r] = orc_normalize(fmaf(om, cfg->dt, th));
                } else { nx[r] = x; ny[r] = y; nth[r] = th; }
            }
            /* a4: provisional footprints of ALL robots, then test each mover (Jacobi form
             * of ConditionalMove/TestCollision, App. A.3) */
            build_grid(g, static_cells, cfg, nx, ny, nth);
            int rebuild = 0;
            for (int r = 0; r < R; ++r) {
                hit[r] = 0;
                if (!moving[r]) continue;
                int n = robot_outline(cfg, nx[r], ny[r], nth[r], ox, oy);
                for (int k = 0; k < n; ++k)
                    if (grid_blocks(g, cfg, ox[k], oy[k], r)) { hit[r] = 1; break; }
            }
            int done_r[ORC_MAX_ROBOTS], live_r[ORC_MAX_ROBOTS];
            for (int r = 0; r < R; ++r) {
                size_t i = base + r;
                float x0 = pose[4 * i + 0], y0 = pose[4 * i + 1], th0 = pose[4 * i + 2];
                if (moving[r]) {
                    if (hit[r]) { nx[r] = x0; ny[r] = y0; nth[r] = th0; meta[4 * i + 2] = 1; rebuild = 1; }
                    else meta[4 * i + 2] = 0;
                }
                int is_live = (live == NULL) || live[i];
                if (cfg->auto_reset == 2 && latch_in[r]) is_live = 0;     /* ppo_stage2.py:72-84 liveflag */
                /* a6: GT velocity by finite difference (stageros.cpp:581-593) */
                float w_gt = orc_normalize(nth[r] - th0) * cfg->inv_dt;
                pose[4 * i + 0] = nx[r]; pose[4 * i + 1] = ny[r]; pose[4 * i + 2] = nth[r];
                float rew; int done = 0, result = 0;
                int crashed = meta[4 * i + 2];
                if (is_live) {
                    /* a10: get_reward_and_terminate (stage_world1.py:180-211) */
                    float ddx = goal[4 * i + 0] - nx[r], ddy = goal[4 * i + 1] - ny[r];
                    float d = sqrtf(fmaf(ddx, ddx, ddy * ddy));
                    float reward_g = (pose[4 * i + 3] - d) * cfg->progress_gain;
                    float reward_c = 0.0f, reward_w = 0.0f;
                    pose[4 * i + 3] = d;
                    if (d < cfg->goal_radius) { done = 1; reward_g = cfg->reward_arrive; result = 1; }
                    if (crashed == 1) { done = 1; reward_c = cfg->reward_collision; result = 2; }
                    if (fabsf(w_gt) > cfg->w_threshold) reward_w = cfg->w_penalty * fabsf(w_gt);
                    if (meta[4 * i + 0] > cfg->timeout) { done = 1; result = 3; }
                    rew = (reward_g + reward_c) + reward_w;
                    acc[4 * i + 0] += rew;
                    acc[4 * i + 1] = rew;
                    meta[4 * i + 0] += 1;
                    meta[4 * i + 3] = done;
                } else {
                    rew = acc[4 * i + 1]; done = 1; result = 0;
                }
                reward[i] = rew;
                flags[4 * i + 0] = (uint8_t)done;
                flags[4 * i + 1] = (uint8_t)crashed;
                flags[4 * i + 2] = (uint8_t)result;
                flags[4 * i + 3] = 0;
                done_r[r] = done; live_r[r] = is_live;
                if (done && is_live) {
                    float *e = eplog + 8 * i;
                    e[0] = goal[4 * i + 0]; e[1] = goal[4 * i + 1]; e[2] = acc[4 * i + 0];
                    e[3] = (float)(meta[4 * i + 0] - 1);
                    e[4] = acc[4 * i + 2]; e[5] = acc[4 * i + 3]; e[6] = (float)result;
                    e[7] = (float)meta[4 * i + 1];
                }
            }
            /* re-spawn: immediately (stage 1, ppo_stage1.py:50-53) or once the whole group has terminated
             * (stage 2: get_group_terminal, model/utils.py:81-87; ppo_stage2.py:105-106) */
            for (int r = 0; r < R; ++r) {
                size_t i = base + r;
                int do_reset = 0;
                if (cfg->auto_reset == 1) do_reset = done_r[r] && live_r[r];
                else if (cfg->auto_reset == 2) {
                    int gid = (int)goal_tab[4 * r + 3];
                    do_reset = 1;
                    for (int r2 = 0; r2 < R; ++r2)
                        if ((int)goal_tab[4 * r2 + 3] == gid && !done_r[r2]) do_reset = 0;
                }
                if (do_reset) {
                    uint32_t gid = (uint32_t)((cfg->world_offset + w) * R + r);
                    reset_agent(cfg, init_tab, goal_tab, gid, r, pose + 4 * i, goal + 4 * i, acc + 4 * i, meta + 4 * i, 0);
                    nx[r] = pose[4 * i + 0]; ny[r] = pose[4 * i + 1]; nth[r] = pose[4 * i + 2];
                    flags[4 * i + 3] = 1;
                    rebuild = 1;
                }
            }
            if (rebuild) build_grid(g, static_cells, cfg, nx, ny, nth);
            /* a5/a8/a9: scan from the final pose, local goal, odom speed */
            for (int r = 0; r < R; ++r) {
                size_t i = base + r;
                scan_robot(g, cfg, r, nx[r], ny[r], nth[r], cosb, sinb, obs + i * cfg->beams, 1);
                float s, c;
                orc_sincosf(nth[r], &s, &c);
                float ddx = goal[4 * i + 0] - nx[r], ddy = goal[4 * i + 1] - ny[r];
                gs[4 * i + 0] = fmaf(ddx, c, ddy * s);            /* get_local_goal, stage_world1.py:155-160 */
                gs[4 * i + 1] = fmaf(ddy, c, -(ddx * s));
                gs[4 * i + 2] = goal[4 * i + 2];                  /* get_self_speed: odom twist = commanded */
                gs[4 * i + 3] = goal[4 * i + 3];
            }
        }
        free(g);
    }
    free(cosb); free(sinb); free(ri);
}

/* Observation only (no tick): the scan / local goal / speed a driver reads right
 * after reset_pose + generate_goal_point (ppo_stage1.py:59-63). */
void orc_observe(const orc_config *cfg, const uint8_t *static_cells,
                 const float *pose, const float *goal, float *obs, float *gs)
{
    int R = cfg->robots_per_world;
    float *cosb = (float *)malloc(sizeof(float) * cfg->beams);
    float *sinb = (float *)malloc(sizeof(float) * cfg->beams);
    int32_t *ri = (int32_t *)malloc(sizeof(int32_t) * cfg->beams);
    orc_beam_table(cfg, cosb, sinb, ri);
#pragma omp parallel
    {
        uint8_t *g = (uint8_t *)malloc((size_t)cfg->grid_w * cfg->grid_h);
        float px[ORC_MAX_ROBOTS], py[ORC_MAX_ROBOTS], pth[ORC_MAX_ROBOTS];
#pragma omp for schedule(dynamic, 1)
        for (int w = 0; w < cfg->num_worlds; ++w) {
            size_t base = (size_t)w * R;
            for (int r = 0; r < R; ++r) {
                px[r] = pose[4 * (base + r) + 0]; py[r] = pose[4 * (base + r) + 1];
                pth[r] = pose[4 * (base + r) + 2];
            }
            build_grid(g, static_cells, cfg, px, py, pth);
            for (int r = 0; r < R; ++r) {
                size_t i = base + r;
                scan_robot(g, cfg, r, px[r], py[r], pth[r], cosb, sinb, obs + i * cfg->beams, 1);
                float s, c;
                orc_sincosf(pth[r], &s, &c);
                float ddx = goal[4 * i + 0] - px[r], ddy = goal[4 * i + 1] - py[r];
                gs[4 * i + 0] = fmaf(ddx, c, ddy * s);
                gs[4 * i + 1] = fmaf(ddy, c, -(ddx * s));
                gs[4 * i + 2] = goal[4 * i + 2];
                gs[4 * i + 3] = goal[4 * i + 3];
            }
        }
        free(g);
    }
    free(cosb); free(sinb); free(ri);
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
