// forwards a ROOT/ROBAST header name to the ROOT-free mirror
#include "../Robast.h"
