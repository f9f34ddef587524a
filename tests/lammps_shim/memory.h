// Forwarding header of the test-only LAMMPS stand-in (see lammps_shim.h).
#pragma once
#include "lammps_shim.h"
