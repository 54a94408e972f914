// Dumps the product's host-side 1-D tables (exadg_b200/csrc/tables.hpp) as text so that the CPU tests can compare
// them with the oracle's independent tables.
#include <cstdio>
#include <cstdlib>

#include "../../exadg_b200/csrc/tables.hpp"

int main(int argc, char ** argv)
{
  const int degree = argc > 1 ? std::atoi(argv[1]) : 3;
  exadg_b200::Tables1D t(degree);
  const int n = t.n;
  auto dump = [&](const char * name, const std::vector<exadg_b200::real_t> & v) {
    std::printf("%s", name);
    for (auto x : v) std::printf(" %.17g", (double)x);
    std::printf("\n");
  };
  std::printf("n %d\n", n);
  dump("xn", t.xn); dump("xq", t.xq); dump("w", t.w); dump("S", t.S); dump("D", t.D); dump("Dq", t.Dq);
  dump("M", t.M); dump("K", t.K); dump("Minv", t.Minv);
  dump("fd0", t.fd[0]); dump("fd1", t.fd[1]); dump("sv0", t.sv[0]); dump("sv1", t.sv[1]); dump("sd0", t.sd[0]); dump("sd1", t.sd[1]);
  return 0;
}
