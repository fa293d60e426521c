// Match.cpp — P/Match.cpp:4-9.
#include "../../../include/Match.hpp"

#include <type_traits>

#include "../../../include/usv_b200.h"

Match::Match(unsigned int LeftIndex, unsigned int RightIndex, double MatchValue) {
  this->LeftIndex = LeftIndex;
  this->RightIndex = RightIndex;
  this->MatchValue = MatchValue;
}

// the GPU writes usv_match records straight into storage that is read back as Match
static_assert(sizeof(Match) == 16 && sizeof(usv_match) == 16, "Match must stay the reference's 16-byte record");
static_assert(std::is_trivially_copyable<Match>::value, "Match must stay trivially copyable");
