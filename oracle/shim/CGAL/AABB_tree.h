#pragma once
#include "basic.h"
