// TEST DRIVER for the checkpoint code of the product's host mirror (experimental-mf_b200/csrc/model.cc):
//   ckpt_tool <mf|dpmf> <train> <nu> <nv> <dim> <model_in> <result_out> <round>
// constructs the model object, init(), read_model() from <model_in>, save_model(<round>) -> "<result_out>_<round>"
// (+ its .state sidecar).  tests/test_checkpoint.py compares the bytes with files written / read by the
// reference's own MF / DPMF::save_model / read_model (oracle/_ref).  Built by experimental-mf_b200/Makefile.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../experimental-mf_b200/csrc/model.h"

int main(int argc, char** argv) {
  if (argc != 9) {
    fprintf(stderr, "usage: ckpt_tool <mf|dpmf> <train> <nu> <nv> <dim> <model_in> <result_out> <round>\n");
    return 2;
  }
  const int nu = atoi(argv[3]), nv = atoi(argv[4]), dim = atoi(argv[5]), round = atoi(argv[8]);
  if (!strcmp(argv[1], "mf")) {
    MF m(argv[2], NULL, argv[7], argv[6], dim, 1, 2e-2f, 1.0f, 5e-3f, 2.76f, nu, nv, 8, 2);
    m.init();
    m.read_model();
    printf("lambda %.9g start_round %d\n", m.lambda_, m.start_round_);
    m.save_model(round);
  } else {
    DPMF m(argv[2], NULL, argv[7], argv[6], dim, 1, 2e-10f, 1.0f, 5e-3f, 2.76f, nu, nv, 8, 2, 1.0f, 100.0f, 0.0f, 0, 1000,
           1.0f, 1e-13f);
    m.init();
    m.read_model();
    printf("lambda_r %.9g start_round %d\n", m.lambda_r_, m.start_round_);
    m.save_model(round);
  }
  return 0;
}
