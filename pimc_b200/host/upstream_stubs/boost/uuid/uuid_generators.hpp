#pragma once
#include "uuid.hpp"
