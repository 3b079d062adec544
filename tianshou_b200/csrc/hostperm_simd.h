// HOST-only helpers of the minibatch-order job (hostperm.cu): the three inner loops of np.random.permutation's stream,
// written so that they vectorise.  Compiled by g++ (hostperm_simd.cpp), no CUDA in here.
//
//   mt_next     one MT19937 state transition (numpy's mt19937_gen, randomkit/_mt19937: the 624-word twist), out of place
//   mt_temper   the tempering of the 624 words of one state
//   walk        the data-dependent part of legacy shuffle's random_interval draws (masked rejection sampling): consumes
//               tempered words, emits the accepted swap partner of every position from the top down
#pragma once
#include <stdint.h>

namespace tsb_hp {

enum Isa { kScalar = 0, kAvx2 = 1, kAvx512 = 2 };

// Best instruction set of this CPU, or what TS_B200_PERM_ISA (scalar | avx2 | avx512) names if the CPU has it.
int detect_isa();

constexpr int kMtN = 624;
// Slack (in words) the `walk` output needs past its last entry: the vector paths store whole registers.
constexpr int kWalkSlack = 64;

void mt_next(const uint32_t* in, uint32_t* out);
void mt_temper(const uint32_t* key, uint32_t* out);

// Walk state: position `i` still to be decided (counts down to 0) and the write cursor of the partner list.  The list is
// in WALK order: entry k belongs to position n - 1 - k.  A rejected candidate is overwritten by the next draw.
struct Walk {
    int64_t i;
    uint32_t* cur;
};
// Consume up to `avail` words; returns how many were consumed (stops early only when i reaches 0).
int walk(int isa, const uint32_t* words, int avail, Walk* w);

}  // namespace tsb_hp
