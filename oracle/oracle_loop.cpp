// oracle_loop.cpp -- command-line front end of the oracle that prints exactly what
// standalone/loop.C:186-195 and standalone/options.h:71-76 print, so that
//   ./oracle_loop | diff /root/reference/standalone/loop.op -
// pins the restatement.  TEST INFRASTRUCTURE ONLY (see oracle.h).
#include "oracle.h"
#include <cstdlib>
#include <cstring>
#include <iostream>

int main(int argc, char** argv) {
  unsigned length = 8, sweeps = 1u << 16;          // options.h:42
  double temperature = 0.2;
  unsigned therm = sweeps >> 3;
  for (int i = 1; i < argc; ++i) {                  // options.h:44-63
    if (!std::strcmp(argv[i], "-l") && i + 1 < argc) length = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "-t") && i + 1 < argc) temperature = std::atof(argv[++i]);
    else if (!std::strcmp(argv[i], "-n") && i + 1 < argc) { sweeps = std::atoi(argv[++i]); therm = sweeps >> 3; }
    else { std::cerr << "usage: oracle_loop [-l int] [-t double] [-n int]\n"; return 1; }
  }
  if (length % 2 == 1 || temperature <= 0. || sweeps == 0) { std::cerr << "invalid parameter\n"; return 1; }
  std::cout << "System Length             = " << length << '\n'
            << "Temperature               = " << temperature << '\n'
            << "MCS for Thermalization    = " << therm << '\n'
            << "MCS for Measurement       = " << sweeps << '\n';
  double o[10];
  orc_run_chain(int(length), temperature, sweeps, therm, o);
  std::cout << "Number of Clusters        = " << o[0] << " +- " << o[1] << std::endl
            << "Energy Density            = " << o[2] << " +- " << o[3] << std::endl
            << "Uniform Susceptibility    = " << o[4] << " +- " << o[5] << std::endl
            << "Staggered Magnetization^2 = " << o[6] << " +- " << o[7] << std::endl
            << "Staggered Susceptibility  = " << o[8] << " +- " << o[9] << std::endl;
}
