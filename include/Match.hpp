// Match.hpp — drop-in for the reference's result record (P/Match.hpp:4-12, P/Match.cpp:4-9):
// same name, same three public fields in the same order, same 3-argument constructor and,
// like the reference, no default constructor. 16 bytes, bit-identical to `usv_match` in
// usv_b200.h, so vectors of Match are filled straight from the GPU's result records.
#ifndef Match_HPP
#define Match_HPP

class Match {
 public:
  Match(unsigned int LeftIndex, unsigned int RightIndex, double MatchValue);  // constructor of the class
  unsigned int LeftIndex;
  unsigned int RightIndex;
  double MatchValue;
};

#endif /* Match_HPP */
