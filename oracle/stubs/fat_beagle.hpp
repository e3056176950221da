// Stand-in for the reference's fat_beagle.hpp in builds without libhmsbeagle (see oracle/Makefile): GPInstance
// only names FatBeagle in MakeLikelihoodTreeEngine / GetLikelihoodTreeEngine (gp_instance.cpp:876-888), which
// are not on the GP path.
#pragma once
#include "phylo_model.hpp"
#include "site_pattern.hpp"
#include "sugar.hpp"
#ifndef BEAGLE_FLAG_VECTOR_SSE
#define BEAGLE_FLAG_VECTOR_SSE 0
#endif
class FatBeagle {
 public:
  FatBeagle(const PhyloModelSpecification&, const SitePattern&, long, bool) {
    Failwith("BEAGLE is not available in this build.");
  }
};
