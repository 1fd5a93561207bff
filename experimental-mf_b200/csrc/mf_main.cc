// ./mf - command-line driver with the reference's flags, defaults, messages and exit codes
// (reference src/main.cc:6-33 help text, :95-137 flag parsing, :138-163 dispatch).  The model
// objects are the B200-backed mirrors of csrc/model.h; the hot path runs in libmf_b200.so.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "model.h"

static void show_help() {  // main.cc:6-33: one line per flag, same names and meaning
  static const char* const lines[] = {
      "--train      xxx       : xxx is the file name of the binary training data.",
      "--nu         int       : number of users.",
      "--nv         int       : number of items.",
      "--test       xxx       : xxx is the file name of the binary test data.",
      "--valid      [xxx]     : xxx is the file name of the binary validation data.",
      "--result     [xxx]     : save your model in name xxx.",
      "--model      [xxx]     : read your model in name xxx.",
      "--alg        [xxx]     : xxx can be {mf, dpmf, admf}.",
      "--dim        [int]     : low rank of the model.",
      "--iter       [int]     : number of iterations.",
      "--fly        [int]     : 1 = single-thread update order; > 1 = parallel schedule (the GPU picks its width).",
      "--stride     [int]     : prefetch strides (accepted, unused on the GPU).",
      "--eta        [float]   : learning rate.",
      "--lambda     [float]   : regularizer.",
      "--gam        [float]   : decay of learning rate.",
      "--bias       [float]   : global bias (important for accuracy).",
      "--mineta     [float]   : minimum learning rate (sometimes used in SGLD).",
      "--epsilon    [float]   : sensitivity of differentially privacy.",
      "--tau        [int]     : maximum of ratings among all the users (usually after trimming your data).",
      "--temp       [float]   : temprature in SGLD (can accelarate the convergence).",
      "--noise_size [int]     : the Gaussian numbers lookup table (accepted, unused: Philox stream).",
      "--eta_reg    [float]   : the learning rate for estimating regularization parameters.",
      "--loss       [int]     : the loss type can be {least square, 0-1 logistic regression}.",
      "--measure    [int]     : 0 = RMSE as the reference computes it; 1 = RMSE after the link of --loss.",
  };
  printf("Usage:\n./mf\n");
  for (const char* l : lines) printf("%s\n", l);
}

int main(int argc, char** argv) {
  // defaults: main.cc:96-105
  char *train_data = NULL, *test_data = NULL, *result = NULL, *alg = NULL, *model = NULL, *valid_data = NULL;
  int dim = 128, iter = 15, tau = 0, nu = 0, nv = 0, fly = 8, stride = 2;
  float eta = 2e-2f, lambda = 5e-3f, gam = 1.0f, mineta = 1e-13f;
  float epsilon = 0.0f, hypera = 1.0f, hyperb = 100.0f, temp = 1.0f;
  float g_bias = 2.76f;
  int noise_size = 2000000000, loss = 0, measure = 0;
  float eta_reg = 2e-3f;

  struct StrFlag { const char* name; char** dst; };
  struct IntFlag { const char* name; int* dst; };
  struct FltFlag { const char* name; float* dst; };
  const StrFlag sf[] = {{"--train", &train_data}, {"--test", &test_data}, {"--valid", &valid_data},
                        {"--result", &result}, {"--model", &model}, {"--alg", &alg}};
  const IntFlag nf[] = {{"--dim", &dim}, {"--iter", &iter}, {"--nu", &nu}, {"--nv", &nv}, {"--fly", &fly},
                        {"--stride", &stride}, {"--tau", &tau}, {"--noise_size", &noise_size},
                        {"--loss", &loss}, {"--measure", &measure}};
  const FltFlag ff[] = {{"--eta", &eta}, {"--lambda", &lambda}, {"--gam", &gam}, {"--bias", &g_bias},
                        {"--mineta", &mineta}, {"--epsilon", &epsilon}, {"--hypera", &hypera},
                        {"--hyperb", &hyperb}, {"--temp", &temp}, {"--eta_reg", &eta_reg}};
  for (int i = 1; i < argc; i++) {
    bool known = false;
    for (const auto& f : sf)
      if (!strcmp(argv[i], f.name)) { if (i + 1 < argc) *f.dst = argv[++i]; known = true; break; }
    if (!known)
      for (const auto& f : nf)
        if (!strcmp(argv[i], f.name)) { if (i + 1 < argc) *f.dst = atoi(argv[++i]); known = true; break; }
    if (!known)
      for (const auto& f : ff)
        if (!strcmp(argv[i], f.name)) { if (i + 1 < argc) *f.dst = (float)atof(argv[++i]); known = true; break; }
    if (!known) {
      printf("%s, unknown parameters, exit\n", argv[i]);  // main.cc:134-135
      exit(1);
    }
  }
  if (train_data == NULL || nu == 0 || nv == 0) {  // main.cc:138-142
    printf("Note that train_data/#users/#items are not optional!\n");
    show_help();
    exit(1);
  }
  // main.cc:143 dereferences alg before testing it for NULL; here a missing --alg means "mf",
  // which is what that line intends
  if (alg == NULL || !strcmp(alg, "mf")) {
    MF mf(train_data, test_data, result, model, dim, iter, eta, gam, lambda, g_bias, nu, nv, fly, stride);
    run(mf);
  } else if (!strcmp(alg, "dpmf")) {
    DPMF dpmf(train_data, test_data, result, model, dim, iter, eta, gam, lambda, g_bias, nu, nv, fly, stride,
              hypera, hyperb, epsilon, tau, noise_size, temp, mineta);
    run(dpmf);
  } else if (!strcmp(alg, "admf")) {
    AdaptRegMF admf(train_data, test_data, valid_data, result, model, dim, iter, eta, gam, lambda, g_bias, nu,
                    nv, fly, stride, loss, measure, eta_reg);
    run(admf);
  } else {
    printf("Pleae select a solver: mf/dpmf/admf\n");  // main.cc:160 (sic)
    exit(2);
  }
  return 0;
}
