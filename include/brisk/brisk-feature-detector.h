// Forwarding header: the drop-in classes live in <brisk/brisk.h> (see that file for reference citations).
#pragma once
#include <brisk/brisk.h>
