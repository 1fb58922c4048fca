// hash_eval.cuh -- the deterministic test evaluator (CRL_EVAL_HASH): policy and value are exact-in-fp32
// functions of a 64-bit hash of the board record, bit-identical to oracle/chessrl_oracle.py
// hash_evaluator().  It stands in for "identical network outputs" when MCTS visit counts are compared
// with the reference's tree code.  Not used for throughput numbers.
#pragma once
#include "chess_core.cuh"

namespace crl {

CRL_HD float hash_policy(u64 h, int label, int bits) {
  u64 r = mix64(h + (u64)label * 0xD6E8FEB86659FD93ULL);
  float scale = 1.0f;
  for (int i = 0; i < bits; ++i) scale *= 0.5f;
  return (float)(r >> (64 - bits)) * scale;
}
CRL_HD float hash_value(u64 h) {
  return (float)(mix64(h ^ 0xA5A5A5A5A5A5A5A5ULL) >> 40) * (1.0f / 8388608.0f) - 1.0f;
}

}  // namespace crl
