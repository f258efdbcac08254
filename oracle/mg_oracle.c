/*
 * oracle/mg_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT THE PRODUCT).  See mg_oracle.h.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference/src/).  "[UPSTREAM]" = public minigrid / stable-baselines3 behaviour as
 * written down in SURVEY.md Appendix A (those packages are not vendored by the reference).
 */
#include "mg_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define MAX_TRIES 1000 /* bound for the reference's `while True` rejection loops (App. B Q7) */
#define ERR_BAD_ACTION 1
#define ERR_TRIES 2

/* ------------------------------------------------------------------ RNG ---------------- */

void mg_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct {
    uint32_t key[2];
    uint32_t ctr[4]; /* ctr[0] = block, ctr[1] = episode, ctr[2..3] = env id */
    uint32_t ndraw;
    uint32_t buf[4];
} rng_t;

static void rng_init(rng_t *r, uint64_t seed, uint64_t env_id, uint32_t episode) {
    r->key[0] = (uint32_t)seed;
    r->key[1] = (uint32_t)(seed >> 32);
    r->ctr[0] = 0;
    r->ctr[1] = episode;
    r->ctr[2] = (uint32_t)env_id;
    r->ctr[3] = (uint32_t)(env_id >> 32);
    r->ndraw = 0;
}

static uint32_t rng_below(rng_t *r, uint32_t n) {
    uint32_t w = r->ndraw & 3u;
    if (w == 0) {
        r->ctr[0] = r->ndraw >> 2;
        mg_philox4x32_10(r->ctr, r->key, r->buf);
    }
    r->ndraw++;
    return (uint32_t)(((uint64_t)r->buf[w] * n) >> 32);
}

/* Python random.randint(a, b): inclusive on both ends */
static int rng_randint(rng_t *r, int a, int b) { return a + (int)rng_below(r, (uint32_t)(b - a + 1)); }

uint32_t mg_draw_below(uint64_t seed, uint64_t env_id, uint32_t episode, uint32_t draw, uint32_t n) {
    rng_t r;
    rng_init(&r, seed, env_id, episode);
    r.ndraw = draw & ~3u;
    uint32_t v = 0;
    for (uint32_t d = r.ndraw; d <= draw; ++d) v = rng_below(&r, n);
    return v;
}

/* ------------------------------------------------------------- kinds / encoding --------- */

static inline int k_is_key(uint8_t k) { return (k >> 3) == 1; }
static inline int k_is_ball(uint8_t k) { return (k >> 3) == 2; }
static inline int k_is_door(uint8_t k) { return k >= MG_K_DOOR && k < MG_K_DOOR + 24; }
static inline int k_is_box(uint8_t k) { return k >= MG_K_BOX; }
static inline int k_colour(uint8_t k) { return k & 7; }
static inline int k_door_state(uint8_t k) { return (k - MG_K_DOOR) >> 3; }

/* [UPSTREAM] WorldObj.encode / Door.encode; empty cell = (1,0,0) (Grid.encode) */
void mg_kind_encode(uint8_t k, uint8_t out[3]) {
    if (k < 8) {
        static const uint8_t t[4] = {1, 2, 8, 9}, c[4] = {0, 5, 1, 0};
        out[0] = t[k & 3]; out[1] = c[k & 3]; out[2] = 0;
    } else if (k_is_key(k)) { out[0] = 5; out[1] = k_colour(k); out[2] = 0; }
    else if (k_is_ball(k)) { out[0] = 6; out[1] = k_colour(k); out[2] = 0; }
    else if (k_is_door(k)) { out[0] = 4; out[1] = k_colour(k); out[2] = (uint8_t)k_door_state(k); }
    else { out[0] = 7; out[1] = k_colour(k); out[2] = 0; }
}

/* [UPSTREAM] MiniGridEnv._reward: 1 - 0.9*(step_count/max_steps) in float64, then the
 * VecEnv float32 reward buffer (README.md:84-95 prints exactly these float32 values). */
void mg_reward_lut(int max_steps, float *lut) {
    for (int k = 0; k <= max_steps; ++k) {
        volatile double q = (double)k / (double)max_steps;
        volatile double m = 0.9 * q;
        volatile double r = 1.0 - m;
        lut[k] = (float)r;
    }
}

/* ---------------------------------------------------------------- generators ------------ */

enum { T_KEY = 0, T_BALL = 1, T_BOX = 2, T_DOOR = 3, T_GOAL = 4 };
/* COLOR_NAMES is sorted alphabetically upstream: blue green grey purple red yellow */
static const uint8_t SORTED_COLOURS[6] = {2, 1, 5, 3, 0, 4};

typedef struct { uint8_t type, colour, x, y; } obj_t;

typedef struct {
    const mg_config *cfg;
    mg_state *s;
    rng_t rng;
    int size, mid;
    obj_t objs[40];
    int nobjs;
    uint8_t pool[24]; /* type*8 + colour, list order of the reference comprehension */
    int npool;
    int agent_x, agent_y; /* (-1,-1) until place_agent */
    int goal_x, goal_y;
} gen_t;

static inline uint8_t *cell(gen_t *g, int x, int y) { return &g->s->grid[y * g->size + x]; }

static void add_obj(gen_t *g, int type, int colour, int x, int y) {
    obj_t o = {(uint8_t)type, (uint8_t)colour, (uint8_t)x, (uint8_t)y};
    g->objs[g->nobjs++] = o;
}

static void pool_fill(gen_t *g, const int *types, int ntypes) {
    g->npool = 0;
    for (int t = 0; t < ntypes; ++t)
        for (int c = 0; c < 6; ++c) g->pool[g->npool++] = (uint8_t)(types[t] * 8 + SORTED_COLOURS[c]);
}

static void pool_remove(gen_t *g, int type, int colour) { /* list.remove((type, colour)) */
    uint8_t e = (uint8_t)(type * 8 + colour);
    for (int i = 0; i < g->npool; ++i)
        if (g->pool[i] == e) {
            memmove(&g->pool[i], &g->pool[i + 1], (size_t)(g->npool - i - 1));
            g->npool--;
            return;
        }
}

static uint8_t pool_take(gen_t *g) { /* choice(pool) then remove */
    int i = (int)rng_below(&g->rng, (uint32_t)g->npool);
    uint8_t e = g->pool[i];
    memmove(&g->pool[i], &g->pool[i + 1], (size_t)(g->npool - i - 1));
    g->npool--;
    return e;
}

static uint8_t obj_kind(int type, int colour) {
    switch (type) {
    case T_KEY: return (uint8_t)(MG_K_KEY + colour);
    case T_BALL: return (uint8_t)(MG_K_BALL + colour);
    case T_BOX: return (uint8_t)(MG_K_BOX + colour);
    case T_DOOR: return (uint8_t)(MG_K_DOOR + 8 + colour); /* Door(colour): closed, unlocked */
    default: return MG_K_GOAL;
    }
}

/* custom_env.py:2036-2046 */
static int next2door(gen_t *g, int x, int y) {
    return k_is_door(*cell(g, x - 1, y)) || k_is_door(*cell(g, x + 1, y)) ||
           k_is_door(*cell(g, x, y - 1)) || k_is_door(*cell(g, x, y + 1));
}

/* [UPSTREAM] MiniGridEnv.place_obj over the whole grid: x drawn first, reject occupied / agent */
static void place_obj(gen_t *g, uint8_t kind, int *px, int *py) {
    int x = 0, y = 0, tries = 0;
    for (;;) {
        x = (int)rng_below(&g->rng, (uint32_t)g->size);
        y = (int)rng_below(&g->rng, (uint32_t)g->size);
        if (++tries >= MAX_TRIES) { g->s->error |= ERR_TRIES; break; }
        if (*cell(g, x, y) != MG_K_EMPTY) continue;
        if (x == g->agent_x && y == g->agent_y) continue;
        break;
    }
    *cell(g, x, y) = kind; /* placing None leaves the cell empty */
    *px = x; *py = y;
}

/* [UPSTREAM] MiniGridEnv.place_agent */
static void place_agent(gen_t *g) {
    int x, y;
    g->agent_x = g->agent_y = -1;
    place_obj(g, MG_K_EMPTY, &x, &y);
    g->agent_x = x; g->agent_y = y;
    g->s->agent_x = (uint8_t)x; g->s->agent_y = (uint8_t)y;
    g->s->agent_dir = (uint8_t)rng_below(&g->rng, 4);
}

/* goal placement, custom_env.py:654-661 / 933-940 / 1394-1401 */
static void place_goal_away_from_doors(gen_t *g) {
    int x, y, tries = 0;
    for (;;) {
        place_obj(g, MG_K_GOAL, &x, &y);
        if (next2door(g, x, y) && ++tries < MAX_TRIES) { *cell(g, x, y) = MG_K_EMPTY; continue; }
        break;
    }
    g->goal_x = x; g->goal_y = y;
    add_obj(g, T_GOAL, 0, x, y);
}

typedef struct { int x0, x1, y0, y1; } room_t;
typedef struct { int colour, locked, key_in_box; } door_t;

/* door colour / locked / key_in_box draws, custom_env.py:635-643, 880-908, 1324-1362 */
static void draw_door_props(gen_t *g, door_t *d, uint8_t *colours, int *ncolours) {
    int i = (int)rng_below(&g->rng, (uint32_t)*ncolours);
    d->colour = colours[i];
    memmove(&colours[i], &colours[i + 1], (size_t)(*ncolours - i - 1));
    (*ncolours)--;
    d->locked = g->cfg->all_doors_open ? 0 : (rng_below(&g->rng, 2) == 0); /* choice([True, False]) */
    d->key_in_box = rng_below(&g->rng, 2) == 0;
    if (d->locked) {
        pool_remove(g, T_KEY, d->colour);
        if (d->key_in_box) pool_remove(g, T_BOX, d->colour);
    }
}

/* door cell, custom_env.py:646-650, 911-929, 1365-1390.  horizontal: on row y=mid at column p */
static void put_door(gen_t *g, const door_t *d, int horizontal, int lo, int hi) {
    int p = rng_randint(&g->rng, lo, hi);
    int is_open = g->cfg->all_doors_open ? (rng_below(&g->rng, 2) == 0) : 0;
    int state = is_open ? 0 : (d->locked ? 2 : 1);
    int x = horizontal ? p : g->mid, y = horizontal ? g->mid : p;
    *cell(g, x, y) = (uint8_t)(MG_K_DOOR + 8 * state + d->colour);
    add_obj(g, T_DOOR, d->colour, x, y);
}

/* key (or box holding it) for a locked door, e.g. custom_env.py:676-693, 958-997, 1461-1478.
 * returns 1 if placed; *kx,*ky = its position */
static int place_key(gen_t *g, const room_t *rm, const door_t *d, int check_agent,
                     int cross_x, int cross_y, int *kx, int *ky) {
    if (!d->locked) return 0;
    int x = 0, y = 0, tries = 0;
    for (;;) {
        x = rng_randint(&g->rng, rm->x0, rm->x1);
        y = rng_randint(&g->rng, rm->y0, rm->y1);
        if (++tries >= MAX_TRIES) { g->s->error |= ERR_TRIES; break; }
        if (x == g->goal_x && y == g->goal_y) continue;
        if (check_agent && x == g->agent_x && y == g->agent_y) continue;
        if (x == cross_x && y == cross_y) continue;
        if (next2door(g, x, y)) continue;
        break;
    }
    if (d->key_in_box) {
        *cell(g, x, y) = (uint8_t)(MG_K_BOX + 8 * (d->colour + 1) + d->colour);
        add_obj(g, T_BOX, d->colour, x, y);
    } else {
        *cell(g, x, y) = (uint8_t)(MG_K_KEY + d->colour);
        add_obj(g, T_KEY, d->colour, x, y);
    }
    *kx = x; *ky = y;
    return 1;
}

/* distractor objects of one room, e.g. custom_env.py:700-725 */
static void place_distractors(gen_t *g, const room_t *rm, int count) {
    for (int n = 0; n < count; ++n) {
        uint8_t e = pool_take(g);
        int type = e >> 3, colour = e & 7;
        int x = 0, y = 0, tries = 0;
        for (;;) {
            x = rng_randint(&g->rng, rm->x0, rm->x1);
            y = rng_randint(&g->rng, rm->y0, rm->y1);
            if (++tries >= MAX_TRIES) { g->s->error |= ERR_TRIES; break; }
            int hit = 0;
            for (int i = 0; i < g->nobjs; ++i)
                if (g->objs[i].x == x && g->objs[i].y == y) { hit = 1; break; }
            if (hit) continue;
            if (x == g->agent_x && y == g->agent_y) continue;
            if (next2door(g, x, y)) continue;
            break;
        }
        *cell(g, x, y) = obj_kind(type, colour);
        add_obj(g, type, colour, x, y);
    }
}

/*
 * _generate_{2,3,4}_rooms (custom_env.py:617-855, 857-1297, 1299-2034) as one table-driven
 * routine.  KEYS[nrooms-2][room][agent_room] lists (up to two) door indices whose key is
 * placed in `room` when the agent starts in `agent_room`; the agent-position check applies
 * only in the agent's own room and the second key avoids the first.
 */
#define NO -1
static const int8_t KEYS[3][4][4][2] = {
    /* 2 rooms: rooms L,R; door 0 */
    {{{0, NO}, {NO, NO}, {NO, NO}, {NO, NO}},
     {{NO, NO}, {0, NO}, {NO, NO}, {NO, NO}},
     {{NO, NO}, {NO, NO}, {NO, NO}, {NO, NO}},
     {{NO, NO}, {NO, NO}, {NO, NO}, {NO, NO}}},
    /* 3 rooms: rooms UL,LL,R; doors h=0, vu=1, vl=2 */
    {{{1, 0}, {NO, NO}, {NO, NO}, {NO, NO}},
     {{NO, NO}, {2, 0}, {NO, NO}, {NO, NO}},
     {{NO, NO}, {NO, NO}, {2, 1}, {NO, NO}},
     {{NO, NO}, {NO, NO}, {NO, NO}, {NO, NO}}},
    /* 4 rooms: rooms UL,LL,UR,LR; doors hl=0, hr=1, vu=2, vl=3 */
    {{{2, 0}, {2, NO}, {0, NO}, {NO, NO}},
     {{3, NO}, {3, 0}, {NO, NO}, {0, NO}},
     {{1, NO}, {NO, NO}, {2, 1}, {2, NO}},
     {{NO, NO}, {1, NO}, {3, NO}, {3, 1}}},
};

static int room_of(int nrooms, int mid, int x, int y) {
    int left = x < mid, upper = y < mid;
    if (nrooms == 2) return left ? 0 : 1;
    if (nrooms == 3) return left ? (upper ? 0 : 1) : 2;
    return left ? (upper ? 0 : 1) : (upper ? 2 : 3);
}

static void generate_rooms(gen_t *g, int nrooms) {
    const int S = g->size, m = g->mid, n = g->cfg->num_objects;
    static const int types[3] = {T_KEY, T_BALL, T_BOX};
    pool_fill(g, types, 3);

    /* walls: x = mid always; y = mid for x<mid (3 rooms) or the full row (4 rooms) */
    for (int i = 1; i < S - 1; ++i) *cell(g, m, i) = MG_K_WALL;
    if (nrooms == 3) for (int i = 1; i < m; ++i) *cell(g, i, m) = MG_K_WALL;
    if (nrooms == 4) for (int i = 1; i < S - 1; ++i) *cell(g, i, m) = MG_K_WALL;

    room_t rooms[4];
    int counts[4];
    int nl = n / 2, nr = n - nl;
    const room_t UL = {1, m - 1, 1, m - 1}, LL = {1, m - 1, m + 1, S - 2};
    const room_t UR = {m + 1, S - 2, 1, m - 1}, LR = {m + 1, S - 2, m + 1, S - 2};
    const room_t L = {1, m - 1, 1, S - 2}, R = {m + 1, S - 2, 1, S - 2};
    if (nrooms == 2) {
        rooms[0] = L; rooms[1] = R; counts[0] = nl; counts[1] = nr;
    } else if (nrooms == 3) {
        rooms[0] = UL; rooms[1] = LL; rooms[2] = R;
        counts[0] = nl / 2; counts[1] = nl - nl / 2; counts[2] = nr;
    } else {
        rooms[0] = UL; rooms[1] = LL; rooms[2] = UR; rooms[3] = LR;
        counts[0] = nl / 2; counts[1] = nl - nl / 2; counts[2] = nr / 2; counts[3] = nr - nr / 2;
    }

    /* door properties first, then door cells, in the reference's order */
    door_t doors[4];
    uint8_t colours[6];
    int ncolours = 6;
    memcpy(colours, SORTED_COLOURS, 6);
    int ndoors = nrooms == 2 ? 1 : nrooms;
    for (int d = 0; d < ndoors; ++d) draw_door_props(g, &doors[d], colours, &ncolours);
    if (nrooms == 2) {
        put_door(g, &doors[0], 0, 1, S - 2);
    } else if (nrooms == 3) {
        put_door(g, &doors[0], 1, 1, m - 1);     /* h  */
        put_door(g, &doors[1], 0, 1, m - 1);     /* vu */
        put_door(g, &doors[2], 0, m + 1, S - 2); /* vl */
    } else {
        put_door(g, &doors[0], 1, 1, m - 1);     /* hl */
        put_door(g, &doors[1], 1, m + 1, S - 2); /* hr */
        put_door(g, &doors[2], 0, 1, m - 1);     /* vu */
        put_door(g, &doors[3], 0, m + 1, S - 2); /* vl */
    }

    place_goal_away_from_doors(g);
    place_agent(g);
    int agent_room = room_of(nrooms, m, g->agent_x, g->agent_y);
    int goal_room = room_of(nrooms, m, g->goal_x, g->goal_y);

    for (int r = 0; r < nrooms; ++r) {
        int kx = -1, ky = -1;
        for (int j = 0; j < 2; ++j) {
            int d = KEYS[nrooms - 2][r][agent_room][j];
            if (d == NO) continue;
            int x, y;
            if (place_key(g, &rooms[r], &doors[d], r == agent_room, kx, ky, &x, &y)) {
                counts[r]--;
                if (j == 0) { kx = x; ky = y; }
            }
        }
        if (goal_room == r) counts[r]--;
        /* Q5 (custom_env.py:1119, 1660): the lower-left loop runs on the upper-left counter */
        int loops = (nrooms >= 3 && r == 1) ? counts[0] : counts[r];
        place_distractors(g, &rooms[r], loops);
    }
}

/* single-room generators, custom_env.py:371-555 */
static void generate_single(gen_t *g, int problem) {
    static const int t_all[4] = {T_KEY, T_BALL, T_BOX, T_DOOR};
    static const int t_gtg[4] = {T_BOX, T_DOOR, T_KEY, T_BALL};
    static const int t_opn[2] = {T_BOX, T_DOOR};
    static const int t_pkp[3] = {T_KEY, T_BOX, T_BALL};
    switch (problem) {
    case MG_P_GTG: pool_fill(g, t_gtg, 4); break;
    case MG_P_OPN: pool_fill(g, t_opn, 2); break;
    case MG_P_PKP: pool_fill(g, t_pkp, 3); break;
    default: pool_fill(g, t_all, 4); break;
    }
    if (problem == MG_P_FULL) { /* _generate_full_map :332-369: every (type, colour), types outer, COLOR_NAMES inner; no draws */
        static const int sorted_colours[6] = {2, 1, 5, 3, 0, 4}; /* blue green grey purple red yellow -> COLOR_TO_IDX */
        for (int t = 0; t < 4; ++t)
            for (int c = 0; c < 6; ++c) {
                int x, y;
                place_obj(g, obj_kind(t_all[t], sorted_colours[c]), &x, &y);
                add_obj(g, t_all[t], sorted_colours[c], x, y);
            }
    } else {
        for (int i = 0; i < g->cfg->num_objects; ++i) {
            uint8_t e = pool_take(g);
            int x, y;
            place_obj(g, obj_kind(e >> 3, e & 7), &x, &y);
            add_obj(g, e >> 3, e & 7, x, y);
        }
    }
    if (problem == MG_P_GTG || problem == MG_P_DRP || problem == MG_P_FULL) {
        int x, y;
        place_obj(g, MG_K_GOAL, &x, &y);
        g->goal_x = x; g->goal_y = y;
        add_obj(g, T_GOAL, 0, x, y);
    }
    place_agent(g);
}

/* obstacles, custom_env.py:155-172 (cfg.reserved = floor((size-2)^2 * percent_obstacles)) */
static void place_obstacles(gen_t *g) {
    const int S = g->size;
    for (int i = 0; i < g->cfg->reserved; ++i) {
        if (g->cfg->problem == MG_P_MULTI) {
            int x = 0, y = 0, tries = 0;
            for (;;) {
                x = rng_randint(&g->rng, 1, S - 2);
                y = rng_randint(&g->rng, 1, S - 2);
                if (++tries >= MAX_TRIES) { g->s->error |= ERR_TRIES; break; }
                if (x == g->mid || y == g->mid) continue;
                int hit = 0;
                for (int k = 0; k < g->nobjs; ++k)
                    if (g->objs[k].x == x && g->objs[k].y == y) { hit = 1; break; }
                if (hit) continue;
                if (x == g->agent_x && y == g->agent_y) continue;
                if (next2door(g, x, y)) continue;
                break;
            }
            *cell(g, x, y) = MG_K_LAVA;
        } else {
            int x, y;
            uint8_t kind = rng_below(&g->rng, 2) == 0 ? MG_K_LAVA : MG_K_WALL; /* choice([Lava(), Wall()]) */
            place_obj(g, kind, &x, &y);
        }
    }
}

static int type4(int t) { return t; } /* key 0 ball 1 box 2 door 3 */

/* _gen_grid, custom_env.py:122-267 */
int mg_generate(const mg_config *cfg, uint64_t seed, uint64_t env_id, mg_state *s) {
    if (cfg->size < 5 || cfg->size > MG_MAX_SIZE) return -1;
    if (cfg->problem < MG_P_MULTI || cfg->problem > MG_P_FULL) return -1;
    gen_t g;
    memset(&g, 0, sizeof g);
    g.cfg = cfg; g.s = s; g.size = cfg->size; g.mid = cfg->size / 2;
    g.agent_x = g.agent_y = -1; g.goal_x = g.goal_y = -1;
    rng_init(&g.rng, seed, env_id, s->episode);

    const int S = g.size;
    memset(s->grid, MG_K_EMPTY, sizeof s->grid);
    for (int i = 0; i < S; ++i) { /* Grid.wall_rect(0,0,w,h) :132 */
        *cell(&g, i, 0) = MG_K_WALL; *cell(&g, i, S - 1) = MG_K_WALL;
        *cell(&g, 0, i) = MG_K_WALL; *cell(&g, S - 1, i) = MG_K_WALL;
    }
    s->carrying = 0; s->step_count = 0;         /* [UPSTREAM] MiniGridEnv.reset */
    s->target_x = s->target_y = MG_NONE; s->target_action = 0; /* :125-127 */
    s->pad = 0;                                                 /* target_range = [] */

    int cmd;
    switch (cfg->problem) {
    case MG_P_MULTI: { /* _generate_multi_map :595-615 */
        static const int cmds[4] = {0, 1, 2, 5};
        cmd = cfg->mission >= 0 ? cfg->mission : cmds[rng_below(&g.rng, 4)];
        generate_rooms(&g, rng_randint(&g.rng, 2, 4));
        break;
    }
    case MG_P_GTO: generate_single(&g, MG_P_GTO); cmd = 0; break;
    case MG_P_GTG: generate_single(&g, MG_P_GTG); cmd = 5; break;
    case MG_P_OPN: generate_single(&g, MG_P_OPN); cmd = 1; break;
    case MG_P_PKP: generate_single(&g, MG_P_PKP); cmd = 2; break;
    case MG_P_DRP: generate_single(&g, MG_P_DRP); cmd = 3; break;
    case MG_P_MOV: generate_single(&g, MG_P_MOV); cmd = 4; break;           /* _generate_move_map :557-593 */
    default: /* _generate_full_map: np_random.choice(msn_commands) after the agent is placed (:365) */
        generate_single(&g, MG_P_FULL);
        cmd = (int)rng_below(&g.rng, 6);
        break;
    }
    if (cfg->reserved > 0) place_obstacles(&g);

    /* target selection :174-267 */
    int tries = 0;
    switch (cmd) {
    case 0: { /* 'go to': np_random.integers(0, len(objs)) until not the goal */
        int i;
        do { i = (int)rng_below(&g.rng, (uint32_t)g.nobjs); }
        while (g.objs[i].type == T_GOAL && ++tries < MAX_TRIES);
        s->mission_id = (uint8_t)(0 * 24 + type4(g.objs[i].type) * 6 + g.objs[i].colour);
        s->target_x = g.objs[i].x; s->target_y = g.objs[i].y; s->target_action = MG_A_DONE;
        break;
    }
    case 1: { /* 'toggle': choice(objs) until box or door */
        int i;
        do { i = (int)rng_below(&g.rng, (uint32_t)g.nobjs); }
        while (!(g.objs[i].type == T_BOX || g.objs[i].type == T_DOOR) && ++tries < MAX_TRIES);
        s->mission_id = (uint8_t)(1 * 24 + type4(g.objs[i].type) * 6 + g.objs[i].colour);
        s->target_x = g.objs[i].x; s->target_y = g.objs[i].y; s->target_action = MG_A_TOGGLE;
        break;
    }
    case 2: { /* 'pick up': choice(objs) until box, key or ball */
        int i;
        do { i = (int)rng_below(&g.rng, (uint32_t)g.nobjs); }
        while (!(g.objs[i].type == T_BOX || g.objs[i].type == T_KEY || g.objs[i].type == T_BALL) &&
               ++tries < MAX_TRIES);
        s->mission_id = (uint8_t)(2 * 24 + type4(g.objs[i].type) * 6 + g.objs[i].colour);
        s->target_x = g.objs[i].x; s->target_y = g.objs[i].y; s->target_action = MG_A_PICKUP;
        break;
    }
    case 3: /* 'drop' :212-214 */
        s->mission_id = MG_MISSION_DROP; s->target_action = MG_A_DROP;
        break;
    case 4: { /* 'move' :216-256: np_random.choice(msn_directions), then the first empty cell of every row / column */
        const int d = (int)rng_below(&g.rng, 4); /* left right up down */
        uint32_t v = 0, p10 = 1;
        for (int k = 1; k <= S - 2; ++k, p10 *= 10) {
            int c = 0;
            if (d == 0) { c = 1; while (c < S - 1 && *cell(&g, c, k) != MG_K_EMPTY) ++c; if (c >= S - 1) c = 0; }
            else if (d == 1) { c = S - 2; while (c > 0 && *cell(&g, c, k) != MG_K_EMPTY) --c; }
            else if (d == 2) { c = 1; while (c < S - 1 && *cell(&g, k, c) != MG_K_EMPTY) ++c; if (c >= S - 1) c = 0; }
            else { c = S - 2; while (c > 0 && *cell(&g, k, c) != MG_K_EMPTY) --c; }
            v += (uint32_t)c * p10;
        }
        s->mission_id = (uint8_t)(MG_MISSION_MOVE0 + d);
        s->target_x = (uint8_t)v; s->target_y = (uint8_t)(v >> 8); s->target_action = (uint8_t)(v >> 16); s->pad = (uint8_t)(v >> 24);
        break;
    }
    case 5: /* 'go to goal' :258-267 */
        s->mission_id = MG_MISSION_GOAL;
        s->target_x = (uint8_t)g.goal_x; s->target_y = (uint8_t)g.goal_y; s->target_action = 0;
        break;
    default:
        return -1;
    }
    if (tries >= MAX_TRIES) s->error |= ERR_TRIES;
    s->reset_draws = (uint16_t)g.rng.ndraw;
    s->episode += 1;
    return (int)g.rng.ndraw;
}

int mg_reset_env(const mg_config *cfg, uint64_t seed, uint64_t env_id, mg_state *s) {
    memset(s, 0, sizeof *s); /* PlaygroundEnv.__init__: mission_done False, reward None (:78-79) */
    return mg_generate(cfg, seed, env_id, s);
}

/* ------------------------------------------------------------------- step --------------- */

static const int DIRX[4] = {1, 0, -1, 0}, DIRY[4] = {0, 1, 0, -1};

/* [UPSTREAM] MiniGridEnv.step followed by PlaygroundEnv.step post-processing (custom_env.py:269-330) */
void mg_step_env(const mg_config *cfg, const float *lut, mg_state *s, int action,
                 float *reward, uint8_t *terminated, uint8_t *truncated, uint8_t *carry_obs) {
    const int S = cfg->size;
    float r = 0.0f;
    int term = 0;

    /* ---- upstream step ---- */
    s->step_count++;
    int fx = s->agent_x + DIRX[s->agent_dir], fy = s->agent_y + DIRY[s->agent_dir];
    uint8_t *fc = &s->grid[fy * S + fx];
    uint8_t k = *fc;
    switch (action) {
    case MG_A_LEFT: s->agent_dir = (uint8_t)((s->agent_dir + 3) & 3); break;
    case MG_A_RIGHT: s->agent_dir = (uint8_t)((s->agent_dir + 1) & 3); break;
    case MG_A_FORWARD:
        if (k == MG_K_EMPTY || k == MG_K_GOAL || k == MG_K_LAVA || (k_is_door(k) && k_door_state(k) == 0)) {
            s->agent_x = (uint8_t)fx; s->agent_y = (uint8_t)fy;
        }
        if (k == MG_K_GOAL) { term = 1; r = lut[s->step_count]; }
        if (k == MG_K_LAVA) term = 1;
        break;
    case MG_A_PICKUP:
        if ((k_is_key(k) || k_is_ball(k) || k_is_box(k)) && s->carrying == 0) { s->carrying = k; *fc = MG_K_EMPTY; }
        break;
    case MG_A_DROP:
        if (k == MG_K_EMPTY && s->carrying != 0) { *fc = s->carrying; s->carrying = 0; }
        break;
    case MG_A_TOGGLE:
        if (k_is_door(k)) {
            int st = k_door_state(k), c = k_colour(k);
            if (st == 2) { /* locked: needs a Key of the same colour */
                if (k_is_key(s->carrying) && k_colour(s->carrying) == c) *fc = (uint8_t)(MG_K_DOOR + c);
            } else {
                *fc = (uint8_t)(MG_K_DOOR + 8 * (st ^ 1) + c);
            }
        } else if (k_is_box(k)) { /* Box.toggle: replaced by its contents */
            int m = (k - MG_K_BOX) >> 3;
            *fc = m ? (uint8_t)(MG_K_KEY + m - 1) : MG_K_EMPTY;
        }
        break;
    case MG_A_DONE: break;
    default: s->error |= ERR_BAD_ACTION; break; /* upstream raises ValueError */
    }
    int trunc = s->step_count >= cfg->max_steps;
    *carry_obs = s->carrying; /* gen_obs() runs here, before the post-processing below (Q2) */

    /* ---- PlaygroundEnv.step post-processing ---- */
    if (term) { /* :272-277 */
        if (s->mission_id != MG_MISSION_GOAL) { s->mission_done = 0; s->latch_step = 0; r = 0.0f; }
        *reward = r; *terminated = 1; *truncated = (uint8_t)trunc;
        return;
    }
    /* front cell of the CURRENT pose (self.front_pos after the move) */
    fx = s->agent_x + DIRX[s->agent_dir]; fy = s->agent_y + DIRY[s->agent_dir];
    if (action == MG_A_TOGGLE) { /* :279-283: colour match only, any object type */
        uint8_t f = s->grid[fy * S + fx];
        if (k_is_door(f) && s->carrying != 0 && k_colour(f) == k_colour(s->carrying)) s->carrying = 0;
    }
    if (!s->mission_done && (unsigned)(s->mission_id - MG_MISSION_MOVE0) < 4u) { /* :314-317: agent_pos in target_range */
        const uint32_t v = (uint32_t)s->target_x | (uint32_t)s->target_y << 8 | (uint32_t)s->target_action << 16 | (uint32_t)s->pad << 24;
        const int d = s->mission_id - MG_MISSION_MOVE0;
        const int k = (d < 2 ? s->agent_y : s->agent_x) - 1, want = d < 2 ? s->agent_x : s->agent_y;
        uint32_t p10 = 1;
        for (int i = 0; i < k; ++i) p10 *= 10;
        if ((int)(v / p10 % 10) == want) { s->mission_done = 1; s->latch_step = s->step_count; }
    } else if (!s->mission_done) { /* :288-312 */
        int arrived = 0;
        if (s->target_x != MG_NONE) {
            if (s->target_action) {
                arrived = (fx == s->target_x && fy == s->target_y); /* :293-297, four direction cases */
            } else if (s->agent_x == s->target_x && s->agent_y == s->target_y) { /* :299-302 */
                s->mission_done = 1; s->latch_step = s->step_count;
            }
        }
        if (arrived && action == s->target_action) { s->mission_done = 1; s->latch_step = s->step_count; }
        if (s->target_x == MG_NONE && s->target_action && action == s->target_action) { /* :309-312 */
            s->mission_done = 1; s->latch_step = s->step_count;
        }
    }
    if (action == MG_A_DONE) { /* :319-328 (manual == False) */
        r = s->mission_done ? lut[s->latch_step] : 0.0f;
        s->mission_done = 0; s->latch_step = 0;
        term = 1;
    }
    *reward = r; *terminated = (uint8_t)term; *truncated = (uint8_t)trunc;
}

/* ------------------------------------------------------------ observations -------------- */

/* [UPSTREAM] gen_obs: view cell (vx,vy) <- world = agent + (vx-3)*right + (6-vy)*dir; outside -> Wall */
void mg_gen_obs(const mg_config *cfg, const mg_state *s, uint8_t carrying, uint8_t *image) {
    const int S = cfg->size;
    const int dx = DIRX[s->agent_dir], dy = DIRY[s->agent_dir];
    const int rx = -dy, ry = dx;
    uint8_t view[MG_VIEW][MG_VIEW];
    uint8_t mask[MG_VIEW][MG_VIEW];
    for (int vx = 0; vx < MG_VIEW; ++vx)
        for (int vy = 0; vy < MG_VIEW; ++vy) {
            int wx = s->agent_x + (vx - 3) * rx + (6 - vy) * dx;
            int wy = s->agent_y + (vx - 3) * ry + (6 - vy) * dy;
            view[vx][vy] = (wx >= 0 && wx < S && wy >= 0 && wy < S) ? s->grid[wy * S + wx] : MG_K_WALL;
            mask[vx][vy] = 1;
        }
    if (!cfg->see_through_walls) { /* Grid.process_vis, agent at (3,6) */
        memset(mask, 0, sizeof mask);
        mask[3][6] = 1;
        for (int j = MG_VIEW - 1; j >= 0; --j) {
            for (int i = 0; i < MG_VIEW - 1; ++i) {
                if (!mask[i][j]) continue;
                uint8_t k = view[i][j];
                if (k == MG_K_WALL || (k_is_door(k) && k_door_state(k) != 0)) continue;
                mask[i + 1][j] = 1;
                if (j > 0) { mask[i + 1][j - 1] = 1; mask[i][j - 1] = 1; }
            }
            for (int i = MG_VIEW - 1; i >= 1; --i) {
                if (!mask[i][j]) continue;
                uint8_t k = view[i][j];
                if (k == MG_K_WALL || (k_is_door(k) && k_door_state(k) != 0)) continue;
                mask[i - 1][j] = 1;
                if (j > 0) { mask[i - 1][j - 1] = 1; mask[i][j - 1] = 1; }
            }
        }
    }
    view[3][6] = carrying; /* the agent's own cell shows what it carries (or empty) */
    for (int vx = 0; vx < MG_VIEW; ++vx)
        for (int vy = 0; vy < MG_VIEW; ++vy) {
            uint8_t *o = &image[(vx * MG_VIEW + vy) * 3];
            if (mask[vx][vy]) mg_kind_encode(view[vx][vy], o);
            else o[0] = o[1] = o[2] = 0;
        }
}

/* [UPSTREAM] FullyObsWrapper.observation (experts_test.py:29) */
void mg_full_obs(const mg_config *cfg, const mg_state *s, uint8_t *image) {
    const int S = cfg->size;
    for (int x = 0; x < S; ++x)
        for (int y = 0; y < S; ++y) mg_kind_encode(s->grid[y * S + x], &image[(x * S + y) * 3]);
    uint8_t *a = &image[(s->agent_x * S + s->agent_y) * 3];
    a[0] = 10; a[1] = 0; a[2] = s->agent_dir;
}

/* ------------------------------------------------------------------ vector env ---------- */

typedef struct {
    const mg_config *cfg; uint64_t seed, base; int lo, hi; mg_state *st; const uint8_t *act;
    uint8_t *obs, *dir, *mis; float *rew; uint8_t *term, *trunc, *ep_len, *term_obs; const float *lut;
    int is_reset;
} job_t;

static void *vec_worker(void *p) {
    job_t *j = (job_t *)p;
    for (int i = j->lo; i < j->hi; ++i) {
        mg_state *s = &j->st[i];
        if (j->is_reset) {
            mg_reset_env(j->cfg, j->seed, j->base + (uint64_t)i, s);
            mg_gen_obs(j->cfg, s, s->carrying, &j->obs[(size_t)i * MG_OBS_BYTES]);
        } else {
            /* [UPSTREAM] DummyVecEnv.step_wait: step; on done keep the terminal obs, reset */
            float r; uint8_t te, tr, co;
            mg_step_env(j->cfg, j->lut, s, j->act[i], &r, &te, &tr, &co);
            j->rew[i] = r; j->term[i] = te; j->trunc[i] = tr;
            if (j->ep_len) j->ep_len[i] = (te | tr) ? s->step_count : 0;
            if (te | tr) {
                if (j->term_obs) mg_gen_obs(j->cfg, s, co, &j->term_obs[(size_t)i * MG_OBS_BYTES]);
                mg_generate(j->cfg, j->seed, j->base + (uint64_t)i, s);
                co = s->carrying;
            }
            mg_gen_obs(j->cfg, s, co, &j->obs[(size_t)i * MG_OBS_BYTES]);
        }
        j->dir[i] = s->agent_dir;
        j->mis[i] = s->mission_id;
    }
    return 0;
}

static void run_jobs(job_t *proto, int n, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads > n) nthreads = n > 0 ? n : 1;
    pthread_t th[256];
    job_t jobs[256];
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = *proto;
        jobs[t].lo = (int)((long long)n * t / nthreads);
        jobs[t].hi = (int)((long long)n * (t + 1) / nthreads);
        if (t > 0) pthread_create(&th[t], 0, vec_worker, &jobs[t]);
    }
    vec_worker(&jobs[0]);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], 0);
}

void mg_vec_reset(const mg_config *cfg, uint64_t seed, uint64_t env_id_base, int n, mg_state *states,
                  uint8_t *obs, uint8_t *dir, uint8_t *mission, int nthreads) {
    job_t j;
    memset(&j, 0, sizeof j);
    j.cfg = cfg; j.seed = seed; j.base = env_id_base; j.st = states;
    j.obs = obs; j.dir = dir; j.mis = mission; j.is_reset = 1;
    run_jobs(&j, n, nthreads);
}

void mg_vec_step(const mg_config *cfg, uint64_t seed, uint64_t env_id_base, int n, mg_state *states,
                 const uint8_t *actions, uint8_t *obs, uint8_t *dir, uint8_t *mission, float *reward,
                 uint8_t *term, uint8_t *trunc, uint8_t *ep_len, uint8_t *term_obs, int nthreads) {
    float lut[MG_GRID_CELLS + 1];
    mg_reward_lut(cfg->max_steps, lut);
    job_t j;
    memset(&j, 0, sizeof j);
    j.cfg = cfg; j.seed = seed; j.base = env_id_base; j.st = states; j.act = actions;
    j.obs = obs; j.dir = dir; j.mis = mission; j.rew = reward; j.term = term; j.trunc = trunc;
    j.ep_len = ep_len; j.term_obs = term_obs; j.lut = lut;
    run_jobs(&j, n, nthreads);
}

/* ------------------------------------------------------------------- GAE ---------------- */

/* [UPSTREAM] SB3 RolloutBuffer.compute_returns_and_advantage: float32 arrays, Python-float
 * gamma/lambda (weak scalars): gamma -> f32; gamma*lambda formed in f64, then -> f32.
 * Order: delta = (r + (g*nv)*nnt) - v;  A = delta + ((gl*nnt)*A).  SURVEY App. A / C-13. */
void mg_gae(const float *rewards, const float *values, const uint8_t *episode_starts,
            const float *last_values, const uint8_t *last_dones, double gamma, double gae_lambda,
            int T, int N, float *advantages, float *returns) {
    const float g = (float)gamma;
    const float gl = (float)(gamma * gae_lambda);
    for (int n = 0; n < N; ++n) {
        volatile float A = 0.0f;
        for (int t = T - 1; t >= 0; --t) {
            float nnt, nv;
            if (t == T - 1) { nnt = 1.0f - (float)last_dones[n]; nv = last_values[n]; }
            else { nnt = 1.0f - (float)episode_starts[(size_t)(t + 1) * N + n]; nv = values[(size_t)(t + 1) * N + n]; }
            volatile float a = g * nv;
            volatile float b = a * nnt;
            volatile float c = rewards[(size_t)t * N + n] + b;
            volatile float delta = c - values[(size_t)t * N + n];
            volatile float d = gl * nnt;
            volatile float e = d * A;
            A = delta + e;
            advantages[(size_t)t * N + n] = A;
            volatile float ret = A + values[(size_t)t * N + n];
            returns[(size_t)t * N + n] = ret;
        }
    }
}
