// Driver for the reference's structure generators (TEST INFRASTRUCTURE - never part of the product path).
//
// The reference programs (screw-dislocation-bcc-fe/screw_dislocation_bcc_fe.cpp, symmetry_tilt_grain_boundary/
// stgb.cpp + stgb_b.cpp) are compiled UNMODIFIED from where they lie: this file includes them with `main` renamed and
// calls their own functions, writing the atoms with full precision (their writers print 6 significant digits).
//   gen_screw <out.txt> [id0 id1 id2]   perfect block; with three atom ids the screw displacement field is applied by
//                                       the reference's screw_dislocation(), which reads the ids from stdin
//   gen_stgb  <out.txt>                 bicrystal of stgb.cpp's default orientation and size
// Output: "n\nLx Ly Lz\n" then n lines "id type x y z" (%.15g).
#include <cfloat>
#include <cstdio>
#include <sstream>
#define main reference_main
#ifdef GEN_SCREW
#include GEN_SRC
#else
#include GEN_SRC
#include GEN_SRC_B
#endif
#undef main

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s <out.txt> [id0 id1 id2]\n", argv[0]); return 2; }
  FILE *fp = fopen(argv[1], "w");
  if (!fp) return 2;
#ifdef GEN_SCREW
  double length_box[3] = {0}, unit_orient[3][3] = {{0}};
  Box::get_length_unitorient(length_box, unit_orient);
  std::vector<struct Coord> coord;
  std::vector<struct Coord>::iterator itc;
  building_matrix(coord, itc, length_box, unit_orient);
  if (argc >= 5) {
    std::istringstream ids(std::string(argv[2]) + " " + argv[3] + " " + argv[4] + "\n");
    std::streambuf *old = std::cin.rdbuf(ids.rdbuf());
    screw_dislocation(coord, itc);
    std::cin.rdbuf(old);
  }
  fprintf(fp, "%zu\n%.15g %.15g %.15g\n", coord.size(), length_box[0], length_box[1], length_box[2]);
  for (itc = coord.begin(); itc != coord.end(); itc++) fprintf(fp, "%d %d %.15g %.15g %.15g\n", itc->id, itc->type, itc->x, itc->y, itc->z);
#else
  std::vector<struct COORD> coord;
  std::vector<struct COORD>::iterator itc;
  double lattice = 2.8553;
  double unit0_xyz[3][3] = {{0.0}};
  double matrix0_xyz[3][3] = {{-1, 1, -2}, {1, -1, -1}, {1, 1, 0}};      // stgb.cpp:21
  double length_box[3] = {34.97014031, 49.45524671, 32.30403188};        // stgb.cpp:22
  for (int i = 0; i < 3; i++) {
    double sum0 = sqrt(pow(matrix0_xyz[i][0], 2) + pow(matrix0_xyz[i][1], 2) + pow(matrix0_xyz[i][2], 2));
    for (int j = 0; j < 3; j++) unit0_xyz[i][j] = matrix0_xyz[i][j] / sum0;
  }
  CRY_BOX matrix0(unit0_xyz, length_box, lattice);                       // the call sequence of stgb.cpp:33-38
  matrix0.get_euler_angle();
  build_crystal(matrix0, coord, itc, 1);
  symm_crystal(matrix0, coord, itc);
  length_box[0] *= 2;
  fprintf(fp, "%zu\n%.15g %.15g %.15g\n", coord.size(), length_box[0], length_box[1], length_box[2]);
  int count = 1;
  for (itc = coord.begin(); itc != coord.end(); itc++) fprintf(fp, "%d %d %.15g %.15g %.15g\n", count++, itc->type, itc->xyz[0], itc->xyz[1], itc->xyz[2]);
#endif
  fclose(fp);
  return 0;
}
