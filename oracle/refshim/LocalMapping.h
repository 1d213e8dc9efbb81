// stand-in for include/LocalMapping.h (included by ProbabilityMapping.cc, nothing used)
#pragma once
