// test stub, see ../stub_common.hpp
#pragma once
#include "../stub_common.hpp"
