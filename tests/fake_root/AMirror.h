#include "fake_root_decls.h"
