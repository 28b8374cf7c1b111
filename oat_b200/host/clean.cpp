// oat-clean NAMES...: remove orphaned "<name>_node" / "<name>_obj" segments (src/cleaner/main.cpp:122-155)
#include <iostream>

#include "oat_host.h"

namespace oat { volatile sig_atomic_t quit = 0; }

int main(int argc, char *argv[])
{
    if (argc < 2) {
        std::cout << "Usage: clean NAMES...\nDeallocate the named shared memory segments specified by NAMES.\n";
        return 0;
    }
    for (int i = 1; i < argc; ++i) {
        const std::string name = argv[i];
        const bool a = oat::Shmem::remove(name + "_node"), b = oat::Shmem::remove(name + "_obj");
        std::cout << "clean: " << ((a || b) ? "Removed" : "Nothing to remove for") << " " << name << "\n";
    }
    return 0;
}
