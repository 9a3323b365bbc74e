/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 * Plain-C restatement of the reference Hanabi environment step + observation encoding:
 *   rules      /root/reference/envs/hanabi/hanabi_lib/hanabi_state.cc, hanabi_hand.cc, hanabi_game.cc
 *   observer   hanabi_observation.cc:52-96
 *   encoder    canonical_encoders.cc:66-109 (hands) 127-171 (board) 192-215 (discards)
 *              240-342 (last action) 370-423 (card knowledge) 465-486 (own hand, fork-added)
 *   env glue   /root/reference/envs/hanabi/rl_env.py:148-267 (reset) 292-442 (step)
 *   deal RNG   std::mt19937 + std::discrete_distribution<unsigned long> (libstdc++ 13,
 *              bits/random.tcc:2657-2714, 3349-3381) as called from hanabi_game.cc:106-112 and
 *              hanabi_state.cc:277-286,313-325.  libstdc++ is a third-party dependency of the
 *              reference, restated here from its published algorithm.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 *
 * Parity status: PINNED.  tests/test_oracle_hanabi.py checks this file against golden episode
 * traces produced by the reference's own Python HanabiEnv (rl_env.py over libpyhanabi.so built
 * unmodified into oracle/_ref; generator tests/golden/make_golden.py) and, live, against
 * oracle/_ref/libref_hanabi.so when present (every observation bit, legal mask, reward, done,
 * score and the full hidden state at every step).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define MAXP 5
#define MAXH 5
#define MAXC 5
#define MAXR 5
enum { MV_PLAY = 1, MV_DISCARD = 2, MV_REVEAL_COLOR = 3, MV_REVEAL_RANK = 4 }; /* hanabi_move.h:33 */

typedef struct {
  /* game (hanabi_game.cc:29-67) */
  int colors, ranks, players, hand_size, max_info, max_life, actions, enc_len, own_len;
  uint32_t mt[624];
  int mti;
  /* state (hanabi_state.h:128-143) */
  int cur_player, next_player, info, life, deck_size, turns_to_play;
  int deck[MAXC * MAXR], discard[MAXC * MAXR], fireworks[MAXC];
  int hand_len[MAXP];
  int card[MAXP][MAXH];  /* colour*ranks+rank */
  int cmask[MAXP][MAXH], rmask[MAXP][MAXH]; /* plausible bitsets (hanabi_hand.h:52-56) */
  int chint[MAXP][MAXH], rhint[MAXP][MAXH]; /* hinted value or -1 */
  /* most recent non-deal history item (hanabi_history_item.h:27-57) */
  int lm_valid, lm_type, lm_player, lm_card_index, lm_target_offset, lm_color, lm_rank;
  int lm_card_color, lm_card_rank, lm_scored, lm_info_token, lm_reveal_bitmask;
} ogame;

/* ---- std::mt19937 ---- */
static void mt_seed(ogame* g, uint32_t s) {
  g->mt[0] = s;
  for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->mti = 624;
}
static uint32_t mt_next(ogame* g) {
  if (g->mti >= 624) {
    for (int k = 0; k < 624; ++k) {
      uint32_t y = (g->mt[k] & 0x80000000u) | (g->mt[(k + 1) % 624] & 0x7fffffffu);
      g->mt[k] = g->mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->mti = 0;
  }
  uint32_t z = g->mt[g->mti++];
  z ^= (z >> 11);
  z ^= (z << 7) & 0x9d2c5680u;
  z ^= (z << 15) & 0xefc60000u;
  z ^= (z >> 18);
  return z;
}

static int card_instances(const ogame* g, int rank) { /* hanabi_game.cc:126-136 */
  if (rank == 0) return 3;
  if (rank == g->ranks - 1) return 1;
  return 2;
}

static int player_to_deal(const ogame* g) { /* hanabi_state.cc:157-164 */
  for (int p = 0; p < g->players; ++p)
    if (g->hand_len[p] < g->hand_size) return p;
  return -1;
}

static void advance(ogame* g) { /* hanabi_state.cc:104-111 */
  if (g->deck_size > 0 && player_to_deal(g) >= 0) {
    g->cur_player = -1;
  } else {
    g->cur_player = g->next_player;
    g->next_player = (g->cur_player + 1) % g->players;
  }
}

/* ApplyRandomChance hanabi_state.cc:282-286 + ChanceOutcomes 313-325 + PickRandomChance + the
 * kDeal branch of ApplyMove 221-243 */
static void deal_random(ogame* g) {
  int n = g->colors * g->ranks;
  int idx[MAXC * MAXR];
  double w[MAXC * MAXR], cp[MAXC * MAXR];
  int m = 0;
  for (int u = 0; u < n; ++u)
    if (g->deck[u] > 0) {
      idx[m] = u;
      w[m] = (double)g->deck[u] / (double)g->deck_size;
      ++m;
    }
  int pick = 0;
  if (m >= 2) { /* discrete_distribution::param_type::_M_initialize */
    double sum = 0.0;
    for (int k = 0; k < m; ++k) sum += w[k];
    double acc = 0.0;
    for (int k = 0; k < m; ++k) {
      double q = w[k] / sum;
      acc = (k == 0) ? q : acc + q;
      cp[k] = acc;
    }
    cp[m - 1] = 1.0;
    /* generate_canonical<double,53>(mt19937): two draws */
    double s = 0.0, tmp = 1.0;
    for (int k = 0; k < 2; ++k) {
      s += (double)mt_next(g) * tmp;
      tmp *= 4294967296.0;
    }
    double p = s / tmp;
    if (p >= 1.0) p = nextafter(1.0, 0.0);
    /* std::lower_bound */
    int lo = 0, len = m;
    while (len > 0) {
      int half = len >> 1;
      if (cp[lo + half] < p) { lo += half + 1; len -= half + 1; }
      else len = half;
    }
    pick = lo;
  }
  int u = idx[pick];
  if (g->deck_size == 0) g->turns_to_play--; /* unreachable: deals need a non-empty deck */
  int to = player_to_deal(g);
  int k = g->hand_len[to]++;
  g->card[to][k] = u;
  g->cmask[to][k] = (1 << g->colors) - 1;
  g->rmask[to][k] = (1 << g->ranks) - 1;
  g->chint[to][k] = -1;
  g->rhint[to][k] = -1;
  g->deck[u]--;
  g->deck_size--;
  advance(g);
}

static int move_is_legal(const ogame* g, int uid) { /* hanabi_state.cc:166-219, uid map hanabi_game.cc:159-183 */
  int h = g->hand_size, P = g->players;
  if (uid < 0 || uid >= g->actions || g->cur_player < 0) return 0;
  if (uid < h) return g->info < g->max_info && uid < g->hand_len[g->cur_player];
  uid -= h;
  if (uid < h) return uid < g->hand_len[g->cur_player];
  uid -= h;
  if (g->info <= 0) return 0;
  if (uid < (P - 1) * g->colors) {
    int off = 1 + uid / g->colors, c = uid % g->colors, t = (g->cur_player + off) % P;
    for (int k = 0; k < g->hand_len[t]; ++k)
      if (g->card[t][k] / g->ranks == c) return 1;
    return 0;
  }
  uid -= (P - 1) * g->colors;
  int off = 1 + uid / g->ranks, r = uid % g->ranks, t = (g->cur_player + off) % P;
  for (int k = 0; k < g->hand_len[t]; ++k)
    if (g->card[t][k] % g->ranks == r) return 1;
  return 0;
}

static void remove_from_hand(ogame* g, int p, int k) { /* hanabi_hand.cc:87-94 */
  for (int j = k; j + 1 < g->hand_len[p]; ++j) {
    g->card[p][j] = g->card[p][j + 1];
    g->cmask[p][j] = g->cmask[p][j + 1];
    g->rmask[p][j] = g->rmask[p][j + 1];
    g->chint[p][j] = g->chint[p][j + 1];
    g->rhint[p][j] = g->rhint[p][j + 1];
  }
  g->hand_len[p]--;
}

static void apply_move(ogame* g, int uid) { /* hanabi_state.cc:221-275 (non-deal branches) */
  int h = g->hand_size, P = g->players, me = g->cur_player;
  if (g->deck_size == 0) g->turns_to_play--;
  g->lm_valid = 1;
  g->lm_player = me;
  g->lm_card_index = -1; g->lm_target_offset = -1; g->lm_color = -1; g->lm_rank = -1;
  g->lm_card_color = -1; g->lm_card_rank = -1; g->lm_scored = 0; g->lm_info_token = 0;
  g->lm_reveal_bitmask = 0;
  if (uid < h) { /* discard */
    g->lm_type = MV_DISCARD;
    g->lm_card_index = uid;
    if (g->info < g->max_info) { g->info++; g->lm_info_token = 1; }
    int c = g->card[me][uid];
    g->lm_card_color = c / g->ranks;
    g->lm_card_rank = c % g->ranks;
    g->discard[c]++;
    remove_from_hand(g, me, uid);
  } else if (uid < 2 * h) { /* play */
    int k = uid - h;
    g->lm_type = MV_PLAY;
    g->lm_card_index = k;
    int c = g->card[me][k], col = c / g->ranks, rk = c % g->ranks;
    g->lm_card_color = col;
    g->lm_card_rank = rk;
    if (rk == g->fireworks[col]) { /* AddToFireworks hanabi_state.cc:132-144 */
      g->fireworks[col]++;
      g->lm_scored = 1;
      if (g->fireworks[col] == g->ranks && g->info < g->max_info) { g->info++; g->lm_info_token = 1; }
    } else {
      g->life--;
      g->discard[c]++;
    }
    remove_from_hand(g, me, k);
  } else if (uid < 2 * h + (P - 1) * g->colors) { /* reveal colour */
    int v = uid - 2 * h, off = 1 + v / g->colors, col = v % g->colors, t = (me + off) % P;
    g->lm_type = MV_REVEAL_COLOR;
    g->lm_target_offset = off;
    g->lm_color = col;
    g->info--;
    for (int k = 0; k < g->hand_len[t]; ++k) { /* HandColorBitmask + RevealColor hanabi_hand.cc:96-110 */
      if (g->card[t][k] / g->ranks == col) {
        g->lm_reveal_bitmask |= 1 << k;
        g->chint[t][k] = col;
        g->cmask[t][k] = 1 << col;
      } else {
        g->cmask[t][k] &= ~(1 << col);
      }
    }
  } else { /* reveal rank */
    int v = uid - 2 * h - (P - 1) * g->colors, off = 1 + v / g->ranks, rk = v % g->ranks, t = (me + off) % P;
    g->lm_type = MV_REVEAL_RANK;
    g->lm_target_offset = off;
    g->lm_rank = rk;
    g->info--;
    for (int k = 0; k < g->hand_len[t]; ++k) {
      if (g->card[t][k] % g->ranks == rk) {
        g->lm_reveal_bitmask |= 1 << k;
        g->rhint[t][k] = rk;
        g->rmask[t][k] = 1 << rk;
      } else {
        g->rmask[t][k] &= ~(1 << rk);
      }
    }
  }
  advance(g);
}

static int score(const ogame* g) { /* hanabi_state.cc:359-364 */
  if (g->life <= 0) return 0;
  int s = 0;
  for (int c = 0; c < g->colors; ++c) s += g->fireworks[c];
  return s;
}

static int is_terminal(const ogame* g) { /* hanabi_state.cc:366-377 */
  if (g->life < 1) return 1;
  if (score(g) >= g->colors * g->ranks) return 1;
  if (g->turns_to_play <= 0) return 1;
  return 0;
}

/* canonical_encoders.cc:441-463: writes enc_len ints (0/1) for observer `obs` */
static void encode(const ogame* g, int obs, int32_t* e) {
  int C = g->colors, R = g->ranks, P = g->players, H = g->hand_size, bpc = C * R;
  int deck_max = 0;
  for (int r = 0; r < R; ++r) deck_max += card_instances(g, r);
  deck_max *= C;
  memset(e, 0, sizeof(int32_t) * g->enc_len);
  int o = 0;
  /* hands :66-109 */
  for (int rel = 1; rel < P; ++rel) {
    int p = (obs + rel) % P;
    for (int k = 0; k < g->hand_len[p]; ++k) e[o + k * bpc + g->card[p][k]] = 1;
    o += H * bpc;
  }
  for (int rel = 0; rel < P; ++rel)
    if (g->hand_len[(obs + rel) % P] < H) e[o + rel] = 1;
  o += P;
  /* board :127-171 */
  for (int i = 0; i < g->deck_size; ++i) e[o + i] = 1;
  o += deck_max - H * P;
  for (int c = 0; c < C; ++c) {
    if (g->fireworks[c] > 0) e[o + g->fireworks[c] - 1] = 1;
    o += R;
  }
  for (int i = 0; i < g->info; ++i) e[o + i] = 1;
  o += g->max_info;
  for (int i = 0; i < g->life; ++i) e[o + i] = 1;
  o += g->max_life;
  /* discards :192-215 */
  for (int c = 0; c < C; ++c)
    for (int r = 0; r < R; ++r) {
      for (int i = 0; i < g->discard[c * R + r]; ++i) e[o + i] = 1;
      o += card_instances(g, r);
    }
  /* last action :240-342; the observation keeps history back to the observer's own previous move
   * (hanabi_observation.cc:80-95), which always contains the most recent non-deal move */
  if (g->lm_valid) {
    int rel = (g->lm_player - obs + P) % P;
    e[o + rel] = 1;
    o += P;
    int ty = g->lm_type;
    e[o + (ty == MV_PLAY ? 0 : ty == MV_DISCARD ? 1 : ty == MV_REVEAL_COLOR ? 2 : 3)] = 1;
    o += 4;
    int reveal = (ty == MV_REVEAL_COLOR || ty == MV_REVEAL_RANK), pd = (ty == MV_PLAY || ty == MV_DISCARD);
    if (reveal) e[o + (rel + g->lm_target_offset) % P] = 1;
    o += P;
    if (ty == MV_REVEAL_COLOR) e[o + g->lm_color] = 1;
    o += C;
    if (ty == MV_REVEAL_RANK) e[o + g->lm_rank] = 1;
    o += R;
    if (reveal)
      for (int i = 0; i < H; ++i)
        if (g->lm_reveal_bitmask & (1 << i)) e[o + i] = 1;
    o += H;
    if (pd) e[o + g->lm_card_index] = 1;
    o += H;
    if (pd) e[o + g->lm_card_color * R + g->lm_card_rank] = 1;
    o += bpc;
    if (ty == MV_PLAY) {
      if (g->lm_scored) e[o] = 1;
      if (g->lm_info_token) e[o + 1] = 1;
    }
    o += 2;
  } else {
    o += P + 4 + P + C + R + H + H + bpc + 2;
  }
  /* card knowledge :370-423 */
  for (int rel = 0; rel < P; ++rel) {
    int p = (obs + rel) % P;
    for (int k = 0; k < g->hand_len[p]; ++k) {
      for (int c = 0; c < C; ++c)
        if (g->cmask[p][k] & (1 << c))
          for (int r = 0; r < R; ++r)
            if (g->rmask[p][k] & (1 << r)) e[o + c * R + r] = 1;
      o += bpc;
      if (g->chint[p][k] >= 0) e[o + g->chint[p][k]] = 1;
      o += C;
      if (g->rhint[p][k] >= 0) e[o + g->rhint[p][k]] = 1;
      o += R;
    }
    o += (H - g->hand_len[p]) * (bpc + C + R);
  }
}

static void encode_own(const ogame* g, int obs, int32_t* e) { /* canonical_encoders.cc:465-486 */
  int bpc = g->colors * g->ranks;
  memset(e, 0, sizeof(int32_t) * g->own_len);
  for (int k = 0; k < g->hand_len[obs]; ++k) e[k * bpc + g->card[obs][k]] = 1;
}

/* rl_env.py:254-263 / 426-434 */
static void observe(const ogame* g, int32_t* out_global, int32_t* out_local, int32_t* out_legal) {
  int cur = g->cur_player, P = g->players;
  if (out_global) {
    encode_own(g, cur, out_global);
    encode(g, cur, out_global + g->own_len);
    for (int p = 0; p < P; ++p) out_global[g->own_len + g->enc_len + p] = (p == cur);
  }
  if (out_local) {
    encode(g, cur, out_local);
    for (int p = 0; p < P; ++p) out_local[g->enc_len + p] = (p == cur);
  }
  if (out_legal)
    for (int a = 0; a < g->actions; ++a) out_legal[a] = move_is_legal(g, a);
}

/* preset 0 = Hanabi-Full, 1 = Hanabi-Small (rl_env.py:110-131) */
void* ohanabi_new(int preset, int seed) {
  ogame* g = (ogame*)calloc(1, sizeof(ogame));
  g->players = 2;
  if (preset == 0) { g->colors = 5; g->ranks = 5; g->hand_size = 5; g->max_info = 8; g->max_life = 3; }
  else { g->colors = 2; g->ranks = 5; g->hand_size = 2; g->max_info = 3; g->max_life = 1; }
  int C = g->colors, R = g->ranks, P = g->players, H = g->hand_size, bpc = C * R;
  int per_color = 0;
  for (int r = 0; r < R; ++r) per_color += card_instances(g, r);
  g->actions = 2 * H + (P - 1) * C + (P - 1) * R;
  g->enc_len = ((P - 1) * H * bpc + P) + (per_color * C - P * H + bpc + g->max_info + g->max_life) +
               per_color * C + (P + 4 + P + C + R + H + H + bpc + 2) + P * H * (bpc + C + R);
  g->own_len = H * bpc;
  mt_seed(g, (uint32_t)seed);
  g->cur_player = -2; /* no state yet */
  return g;
}

void ohanabi_free(void* h) { free(h); }

void ohanabi_dims(void* h, int* out) {
  ogame* g = (ogame*)h;
  out[0] = g->enc_len; out[1] = g->own_len; out[2] = g->players; out[3] = g->actions;
  out[4] = g->colors; out[5] = g->ranks; out[6] = g->hand_size; out[7] = g->max_info; out[8] = g->max_life;
}

static void new_state(ogame* g) { /* HanabiState ctor hanabi_state.cc:90-102, HanabiDeck 53-64 */
  g->deck_size = 0;
  for (int c = 0; c < g->colors; ++c)
    for (int r = 0; r < g->ranks; ++r) {
      g->deck[c * g->ranks + r] = card_instances(g, r);
      g->discard[c * g->ranks + r] = 0;
      g->deck_size += card_instances(g, r);
    }
  for (int p = 0; p < g->players; ++p) g->hand_len[p] = 0;
  for (int c = 0; c < g->colors; ++c) g->fireworks[c] = 0;
  g->cur_player = -1;
  g->next_player = 0; /* GetSampledStartPlayer with random_start_player=false */
  g->info = g->max_info;
  g->life = g->max_life;
  g->turns_to_play = g->players;
  g->lm_valid = 0;
}

void ohanabi_reset(void* h, int32_t* out_global, int32_t* out_local, int32_t* out_legal) {
  ogame* g = (ogame*)h;
  new_state(g);
  while (g->cur_player == -1) deal_random(g);
  observe(g, out_global, out_local, out_legal);
}

/* returns 0, or -1 if the action is illegal (the reference aborts: REQUIRE, hanabi_state.cc:222) */
int ohanabi_step(void* h, int action, int32_t* out_global, int32_t* out_local, int32_t* out_legal,
                 int32_t* out_rds) {
  ogame* g = (ogame*)h;
  if (!move_is_legal(g, action)) return -1;
  int last = score(g);
  apply_move(g, action);
  while (g->cur_player == -1) deal_random(g);
  observe(g, out_global, out_local, out_legal);
  out_rds[0] = score(g) - last;
  out_rds[1] = is_terminal(g);
  out_rds[2] = score(g);
  return 0;
}

/* same layout as ref_env_dump (oracle/ref_wrap/hanabi_capi.cpp) */
int ohanabi_dump(void* h, int32_t* out) {
  ogame* g = (ogame*)h;
  int C = g->colors, R = g->ranks, H = g->hand_size, o = 0;
  out[o++] = g->cur_player; out[o++] = g->info; out[o++] = g->life; out[o++] = g->deck_size;
  out[o++] = is_terminal(g);
  for (int c = 0; c < C; ++c) out[o++] = g->fireworks[c];
  for (int i = 0; i < C * R; ++i) out[o++] = g->deck[i];
  for (int i = 0; i < C * R; ++i) out[o++] = g->discard[i];
  for (int p = 0; p < g->players; ++p) {
    out[o++] = g->hand_len[p];
    for (int k = 0; k < H; ++k) {
      if (k < g->hand_len[p]) {
        out[o++] = g->card[p][k]; out[o++] = g->cmask[p][k]; out[o++] = g->rmask[p][k];
        out[o++] = g->chint[p][k]; out[o++] = g->rhint[p][k];
      } else {
        out[o++] = -1; out[o++] = 0; out[o++] = 0; out[o++] = -1; out[o++] = -1;
      }
    }
  }
  return o;
}

/* throughput harness: same policy/LCG and same per-step work (both players observed + encoded)
 * as ref_env_play in oracle/ref_wrap/hanabi_capi.cpp */
long ohanabi_play(void* h, long steps, unsigned lcg_seed, long* out_checksum) {
  ogame* g = (ogame*)h;
  unsigned long long lcg = lcg_seed * 2862933555777941757ULL + 3037000493ULL;
  long sum = 0;
  int32_t v[1024], own[128];
  if (g->cur_player == -2) { new_state(g); while (g->cur_player == -1) deal_random(g); }
  for (long t = 0; t < steps; ++t) {
    int legal[64], nl = 0;
    for (int a = 0; a < g->actions; ++a) if (move_is_legal(g, a)) legal[nl++] = a;
    lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
    apply_move(g, legal[(lcg >> 33) % (unsigned)nl]);
    while (g->cur_player == -1) deal_random(g);
    for (int p = 0; p < g->players; ++p) {
      encode(g, p, v);
      encode_own(g, p, own);
      int n_legal = 0;
      if (p == g->cur_player) for (int a = 0; a < g->actions; ++a) n_legal += move_is_legal(g, a);
      sum += v[t % g->enc_len] + own[t % g->own_len] + n_legal;
    }
    if (is_terminal(g)) { new_state(g); while (g->cur_player == -1) deal_random(g); }
  }
  if (out_checksum) *out_checksum = sum;
  return steps;
}
