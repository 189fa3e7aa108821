/* pgshim: everything lives in postgres.h */
#include "postgres.h"
