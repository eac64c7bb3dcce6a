// ORACLE shim (test infrastructure): see boost/shim_common.hpp
#pragma once
#include <boost/shim_common.hpp>
