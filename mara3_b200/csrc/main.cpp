/**
 * main.cpp -- `mara3b <subprogram> key=value ...`: the command line of the reference's `mara` executable
 * (Mara3 src/app_main.cpp:41-79) for the one subprogram this library implements, `binary`.
 * Everything happens behind the C ABI (include/mara3_b200.h: m3b_binary_main).
 */
#include <cstdlib>
#include <cstring>
#include <iostream>
#include "../../include/mara3_b200.h"

int main(int argc, const char* argv[])
{
    if (argc == 1)
    {
        std::cout << "usages: \n    mara3b binary" << std::endl;
        return 0;
    }
    if (! std::strcmp(argv[1], "binary"))
    {
        const char* dev = std::getenv("M3B_DEVICE");
        return m3b_binary_main(argc - 1, argv + 1, dev ? std::atoi(dev) : 0);
    }
    std::cout << "invalid sub-program '" << argv[1] << "'\n";
    return 0;
}
