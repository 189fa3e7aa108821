#include "../postgres.h"
