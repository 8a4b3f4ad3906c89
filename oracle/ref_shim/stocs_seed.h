// TEST INFRASTRUCTURE ONLY: the engine seed SelectQuadrilateralStoCS uses in the oracle build (see oracle/Makefile, second patch).
#pragma once
extern "C" unsigned pgp_oracle_stocs_seed;
