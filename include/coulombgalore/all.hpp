// coulombgalore/all.hpp -- umbrella include (reference all.hpp:27-38).  The JSON factory createScheme (all.hpp:42-104)
// is configuration plumbing outside the pair path and is not provided; schemes are constructed directly.
#pragma once
#include "batched.hpp"
#include "ewald.hpp"
#include "ewald_truncated.hpp"
#include "fanourgakis.hpp"
#include "fennell.hpp"
#include "plain.hpp"
#include "poisson.hpp"
#include "qpotential.hpp"
#include "reactionfield.hpp"
#include "splined.hpp"
#include "wolf.hpp"
#include "zahn.hpp"
#include "zerodipole.hpp"
