// TEST INFRASTRUCTURE (oracle/_ref). A plain C ABI around the UNMODIFIED reference Hanabi
// library (hanabi_lib/*.cc compiled from where they lie under /root/reference/envs/hanabi).
// It drives the reference classes in the order HanabiEnv.reset/step does
// (/root/reference/envs/hanabi/rl_env.py:148-267, 292-442):
//   step : get_move(uid) -> score() -> ApplyMove -> deal while chance -> observation(cur_p)
//          -> Encode + EncodeOwnHand -> legal uids -> IsTerminal -> reward = score - last_score
//   reset: new HanabiState(game) (the HanabiGame and its mt19937 persist) -> deal until full
// No game logic is re-implemented here: all transitions, legality, encoding and RNG draws are
// the reference's own (HanabiState / HanabiObservation / CanonicalObservationEncoder /
// HanabiGame::PickRandomChance).  Used (a) to pin the C restatement in oracle/hanabi_oracle.c,
// (b) to generate tests/golden fixtures, (c) as the "reference" CPU baseline in bench.py.
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "canonical_encoders.h"
#include "hanabi_game.h"
#include "hanabi_observation.h"
#include "hanabi_state.h"

using namespace hanabi_learning_env;

namespace {
struct RefEnv {
  HanabiGame* game = nullptr;
  HanabiState* state = nullptr;
  CanonicalObservationEncoder* enc = nullptr;
  int enc_len = 0, own_len = 0, players = 0, actions = 0;
};

// presets of rl_env.py:110-131
std::unordered_map<std::string, std::string> preset_params(int preset, int seed) {
  std::unordered_map<std::string, std::string> p;
  if (preset == 0) {  // Hanabi-Full
    p = {{"colors", "5"}, {"ranks", "5"}, {"players", "2"}, {"max_information_tokens", "8"},
         {"max_life_tokens", "3"}, {"observation_type", "1"}};
  } else {  // Hanabi-Small
    p = {{"colors", "2"}, {"ranks", "5"}, {"players", "2"}, {"hand_size", "2"},
         {"max_information_tokens", "3"}, {"max_life_tokens", "1"}, {"observation_type", "1"}};
  }
  p["seed"] = std::to_string(seed);
  return p;
}

// writes: global = own-hand ‖ encode ‖ turn one-hot; local = encode ‖ turn one-hot; legal mask
void observe(RefEnv* e, int32_t* out_global, int32_t* out_local, int32_t* out_legal) {
  int cur = e->state->CurPlayer();
  HanabiObservation obs(*e->state, cur);
  std::vector<int> v = e->enc->Encode(obs);
  std::vector<int> own = e->enc->EncodeOwnHand(obs);
  if (out_global) {
    int o = 0;
    for (int x : own) out_global[o++] = x;
    for (int x : v) out_global[o++] = x;
    for (int p = 0; p < e->players; ++p) out_global[o++] = (p == cur);
  }
  if (out_local) {
    int o = 0;
    for (int x : v) out_local[o++] = x;
    for (int p = 0; p < e->players; ++p) out_local[o++] = (p == cur);
  }
  if (out_legal) {
    for (int a = 0; a < e->actions; ++a) out_legal[a] = 0;
    for (const HanabiMove& m : obs.LegalMoves()) out_legal[e->game->GetMoveUid(m)] = 1;
  }
}
}  // namespace

extern "C" {

void* ref_env_new(int preset, int seed) {
  RefEnv* e = new RefEnv;
  e->game = new HanabiGame(preset_params(preset, seed));
  e->enc = new CanonicalObservationEncoder(e->game);
  e->enc_len = e->enc->Shape()[0];
  e->own_len = e->enc->OwnHandShape()[0];
  e->players = e->game->NumPlayers();
  e->actions = e->game->MaxMoves();
  return e;
}

void ref_env_free(void* h) {
  RefEnv* e = (RefEnv*)h;
  delete e->state;
  delete e->enc;
  delete e->game;
  delete e;
}

// dims: [enc_len, own_len, players, actions, colors, ranks, hand_size, max_info, max_life]
void ref_env_dims(void* h, int* out) {
  RefEnv* e = (RefEnv*)h;
  out[0] = e->enc_len;
  out[1] = e->own_len;
  out[2] = e->players;
  out[3] = e->actions;
  out[4] = e->game->NumColors();
  out[5] = e->game->NumRanks();
  out[6] = e->game->HandSize();
  out[7] = e->game->MaxInformationTokens();
  out[8] = e->game->MaxLifeTokens();
}

void ref_env_reset(void* h, int32_t* out_global, int32_t* out_local, int32_t* out_legal) {
  RefEnv* e = (RefEnv*)h;
  delete e->state;
  e->state = new HanabiState(e->game);
  while (e->state->CurPlayer() == kChancePlayerId) e->state->ApplyRandomChance();
  observe(e, out_global, out_local, out_legal);
}

// out_rds = [reward, done, score]
void ref_env_step(void* h, int action, int32_t* out_global, int32_t* out_local, int32_t* out_legal,
                  int32_t* out_rds) {
  RefEnv* e = (RefEnv*)h;
  HanabiMove move = e->game->GetMove(action);
  int last_score = e->state->Score();
  e->state->ApplyMove(move);
  while (e->state->CurPlayer() == kChancePlayerId) e->state->ApplyRandomChance();
  observe(e, out_global, out_local, out_legal);
  out_rds[0] = e->state->Score() - last_score;
  out_rds[1] = e->state->IsTerminal() ? 1 : 0;
  out_rds[2] = e->state->Score();
}

// Full hidden state dump, layout shared with oracle/hanabi_oracle.c (hz_state_dump):
//  [0] cur_player [1] info [2] life [3] deck_size [4] turns-proxy (see below) [5..5+C) fireworks
//  then deck counts[C*R], discard counts[C*R], then per player: hand_len, per slot (H slots):
//  card index (c*R+r or -1), colour-plausible mask, rank-plausible mask, hinted colour, hinted rank.
// turns_to_play_ is private in the reference, so slot [4] carries IsTerminal() instead.
int ref_env_dump(void* h, int32_t* out) {
  RefEnv* e = (RefEnv*)h;
  const HanabiState& s = *e->state;
  int C = e->game->NumColors(), R = e->game->NumRanks(), H = e->game->HandSize();
  int o = 0;
  out[o++] = s.CurPlayer();
  out[o++] = s.InformationTokens();
  out[o++] = s.LifeTokens();
  out[o++] = s.Deck().Size();
  out[o++] = s.IsTerminal() ? 1 : 0;
  for (int c = 0; c < C; ++c) out[o++] = s.Fireworks()[c];
  for (int c = 0; c < C; ++c)
    for (int r = 0; r < R; ++r) out[o++] = s.Deck().CardCount(c, r);
  std::vector<int> disc(C * R, 0);
  for (const HanabiCard& card : s.DiscardPile()) ++disc[card.Color() * R + card.Rank()];
  for (int i = 0; i < C * R; ++i) out[o++] = disc[i];
  for (int p = 0; p < e->players; ++p) {
    const HanabiHand& hand = s.Hands()[p];
    int n = (int)hand.Cards().size();
    out[o++] = n;
    for (int k = 0; k < H; ++k) {
      if (k < n) {
        const auto& kn = hand.Knowledge()[k];
        int cm = 0, rm = 0;
        for (int c = 0; c < C; ++c) cm |= (kn.ColorPlausible(c) ? 1 : 0) << c;
        for (int r = 0; r < R; ++r) rm |= (kn.RankPlausible(r) ? 1 : 0) << r;
        out[o++] = hand.Cards()[k].Color() * R + hand.Cards()[k].Rank();
        out[o++] = cm;
        out[o++] = rm;
        out[o++] = kn.Color();
        out[o++] = kn.Rank();
      } else {
        out[o++] = -1; out[o++] = 0; out[o++] = 0; out[o++] = -1; out[o++] = -1;
      }
    }
  }
  return o;
}

// Throughput harness for the CPU baseline: plays `steps` env steps with the policy
// "uniformly random legal move from a private LCG", auto-resetting finished games, doing
// exactly the per-step work rl_env.step asks of the C++ side for BOTH players
// (_make_observation_all_players, rl_env.py:443-455): observation + Encode + EncodeOwnHand.
// Returns the number of steps played; checksum defeats dead-code elimination.
long ref_env_play(void* h, long steps, unsigned lcg_seed, long* out_checksum) {
  RefEnv* e = (RefEnv*)h;
  unsigned long long lcg = lcg_seed * 2862933555777941757ULL + 3037000493ULL;
  long sum = 0;
  if (!e->state) {
    e->state = new HanabiState(e->game);
    while (e->state->CurPlayer() == kChancePlayerId) e->state->ApplyRandomChance();
  }
  for (long t = 0; t < steps; ++t) {
    std::vector<HanabiMove> legal = e->state->LegalMoves(e->state->CurPlayer());
    lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
    HanabiMove move = legal[(lcg >> 33) % legal.size()];
    e->state->ApplyMove(move);
    while (e->state->CurPlayer() == kChancePlayerId) e->state->ApplyRandomChance();
    for (int p = 0; p < e->players; ++p) {
      HanabiObservation obs(*e->state, p);
      std::vector<int> v = e->enc->Encode(obs);
      std::vector<int> own = e->enc->EncodeOwnHand(obs);
      sum += v[(size_t)(t % v.size())] + own[(size_t)(t % own.size())] + (long)obs.LegalMoves().size();
    }
    if (e->state->IsTerminal()) {
      delete e->state;
      e->state = new HanabiState(e->game);
      while (e->state->CurPlayer() == kChancePlayerId) e->state->ApplyRandomChance();
    }
  }
  if (out_checksum) *out_checksum = sum;
  return steps;
}

}  // extern "C"
