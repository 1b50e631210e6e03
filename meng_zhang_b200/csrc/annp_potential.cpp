// annp_potential.cpp -- reader for the `.ann` potential format (host only, no CUDA).
//
// Reproduces the parsing contract of PairANNP::read_file (reference
// annp-gpu-lammps/fe_v2/src/pair_annp.cpp:332-518) so that the same files give the same numbers:
//   * records are addressed by absolute line index (line k, 1-based, is loop index k-1)
//   * a number is taken with atof at column 0 and after every TAB whose next character is a digit
//     or '-' (so ".5" or "+1" after a tab would be skipped, exactly as in the reference)
//   * CRLF line ends are tolerated (atof stops at '\r')
//   * activation / descriptor keywords are found by two-character scans: "Ch"->Chebyshev,
//     "Be"/"BP"->1, "Cu"->2; "li"->0, "hy"->1, "si"->2, "mo"->3, "ta"->4  (so "tanh" selects 4)
//   * weight blocks `#<layer>_(weight|bias)` are always stored under element 0: the reference resets
//     its element index on every line it reads (pair_annp.cpp:455), which makes the format
//     effectively single-element
#include "../../include/annp_b200.h"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

namespace {

void set_err(char *err, int errlen, const std::string &msg) {
  if (err && errlen > 0) {
    snprintf(err, (size_t) errlen, "%s", msg.c_str());
  }
}

bool tab_number(const std::string &s, size_t j, bool allow_minus) {
  if (s[j] != '\t' || j + 1 >= s.size()) return false;
  const unsigned char nx = (unsigned char) s[j + 1];
  return isdigit(nx) || (allow_minus && nx == '-');
}

// values of one row: atof at 0, then after each qualifying tab
std::vector<double> row_values(const std::string &s, bool allow_minus) {
  std::vector<double> v;
  v.push_back(atof(s.c_str()));
  for (size_t j = 0; j < s.size(); j++)
    if (tab_number(s, j, allow_minus)) v.push_back(atof(s.c_str() + j + 1));
  return v;
}

}    // namespace

extern "C" size_t annp_b200_weights_per_element(int ntl, int nnod, int nsf) {
  const int nl = ntl - 1;
  if (nl <= 0) return 0;
  if (nl == 1) return (size_t) nsf;
  return (size_t) nnod * nsf + (size_t) (nl - 2) * nnod * nnod + (size_t) nnod;
}

extern "C" size_t annp_b200_bias_per_element(int ntl, int nnod) {
  const int nl = ntl - 1;
  if (nl <= 0) return 0;
  return (size_t) (nl - 1) * nnod + 1;
}

extern "C" void annp_b200_free_potential(annp_b200_potential *pot) {
  if (!pot) return;
  free(pot->weight_all);
  free(pot->bias_all);
  pot->weight_all = nullptr;
  pot->bias_all = nullptr;
}

extern "C" int annp_b200_read_potential(const char *filename, int nelements_coeff, const char *const *elements_coeff,
                                        annp_b200_potential *out, char *err, int errlen) {
  (void) elements_coeff;
  if (!filename || !out) { set_err(err, errlen, "null argument"); return ANNP_B200_EINVAL; }
  memset(out, 0, sizeof(*out));
  std::ifstream fin(filename, std::ios::in);
  if (!fin.is_open()) {
    set_err(err, errlen, "Cannot open neural network potential file");   // pair_annp.cpp:341
    return ANNP_B200_EIO;
  }
  std::string t;
  int nel = 0;
  for (int i = 0; i < 21 + nelements_coeff; i++) {
    if (!std::getline(fin, t)) { set_err(err, errlen, "potential file truncated in header"); return ANNP_B200_EIO; }
    if (i == 5) {
      nel = out->nelements = atoi(t.c_str());
      if (nel < 1 || nel > ANNP_B200_MAX_ELEMENTS) { set_err(err, errlen, "unsupported number of elements"); return ANNP_B200_EINVAL; }
    }
    if (i >= 6 && i < 6 + nel) {
      const int e = i - 6;
      out->id_elem[e] = atoi(t.c_str());
      std::string name;
      for (size_t j = 0; j < t.size(); j++) {
        if (isalpha((unsigned char) t[j])) name += t[j];
        if (tab_number(t, j, false)) out->mass[e] = atof(t.c_str() + j + 1);
      }
      snprintf(out->elements[e], sizeof(out->elements[e]), "%s", name.c_str());
    }
    if (i == 8 + nel) {
      out->ntl = atoi(t.c_str());
      int np = 1;
      for (size_t j = 0; j < t.size(); j++) {
        if (tab_number(t, j, false)) {
          const char *p = t.c_str() + j + 1;
          if (np == 1) out->nhl = atoi(p);
          if (np == 2) out->nnod = atoi(p);
          if (np == 3) out->nsf = atoi(p);
          if (np == 4) out->npsf = atoi(p);
          if (np == 5) out->ntsf = atoi(p);
          if (np == 6) out->cut = atof(p);
          np++;
        }
      }
      if (out->ntl < 2 || out->ntl - 1 > ANNP_B200_MAX_LAYERS || out->nnod < 1 || out->nnod > ANNP_B200_MAX_NOD ||
          out->nsf < 1 || out->nsf > ANNP_B200_MAX_SF) {
        set_err(err, errlen, "network dimensions outside the supported range");
        return ANNP_B200_EINVAL;
      }
    }
    if (i >= 11 + nel && i <= 12 + nel) {
      std::vector<double> v = row_values(t, true);
      if ((int) v.size() > ANNP_B200_MAX_SF) { set_err(err, errlen, "too many normalisation values"); return ANNP_B200_EINVAL; }
      double *dst = (i == 11 + nel) ? out->sfnor_cov : out->sfnor_avg;
      for (size_t k = 0; k < v.size(); k++) dst[k] = v[k];
    }
    if (i == 15 + nel) {
      int nact = 0;
      const int nlayer = out->ntl - 1;
      for (size_t j = 0; j + 1 < t.size() + 1; j++) {
        const char a = t[j], b = (j + 1 < t.size()) ? t[j + 1] : '\0';
        if (a == 'C' && b == 'h') out->flagsym = 0;
        if ((a == 'B' && b == 'e') || (a == 'B' && b == 'P')) out->flagsym = 1;
        if (a == 'C' && b == 'u') out->flagsym = 2;
        int act = -1;
        if (a == 'l' && b == 'i') act = 0;
        if (a == 'h' && b == 'y') act = 1;
        if (a == 's' && b == 'i') act = 2;
        if (a == 'm' && b == 'o') act = 3;
        if (a == 't' && b == 'a') act = 4;
        if (act >= 0) {
          if (nact >= nlayer || nact >= ANNP_B200_MAX_LAYERS) { set_err(err, errlen, "more activation keywords than layers"); return ANNP_B200_EINVAL; }
          out->flagact[nact++] = act;
        }
      }
    }
    if (i == 18 + nel) out->e_scale = atof(t.c_str());
    if (i == 19 + nel) out->e_shift = atof(t.c_str());
    if (i == 20 + nel) out->e_atom = atof(t.c_str());
  }
  if (out->ntl < 2) { set_err(err, errlen, "potential file has no network record"); return ANNP_B200_EIO; }

  const int n_lay = out->ntl - 1, n_nod = out->nnod, n_sf = out->nsf;
  out->weight_all = (double *) calloc((size_t) nel * n_lay * n_nod * n_sf, sizeof(double));
  out->bias_all = (double *) calloc((size_t) nel * n_lay * n_nod, sizeof(double));
  if (!out->weight_all || !out->bias_all) { annp_b200_free_potential(out); set_err(err, errlen, "out of host memory"); return ANNP_B200_ENOMEM; }

  while (true) {
    if (!std::getline(fin, t)) break;
    const int type_elem = 0;   // see header comment
    if (t.size() >= 2 && t[0] == '#' && isdigit((unsigned char) t[1])) {
      int no_layer = 0;
      bool flag_wb = false;
      for (size_t i = 0; i < t.size(); i++) {
        if (t[i] > 47 && t[i] < 58) { no_layer *= 10; no_layer += t[i] - 48; }
        if (t[i] == 'w') flag_wb = false;
        if (t[i] == 'b') flag_wb = true;
      }
      int nrow_w = n_nod, ncol_w = n_nod, nrow_b = 1, ncol_b = n_nod;
      if (no_layer == 1) { nrow_w = n_nod; ncol_w = n_sf; }
      if (no_layer == out->ntl - 1) { nrow_w = 1; ncol_w = n_nod; nrow_b = 1; ncol_b = 1; }
      (void) ncol_b;
      const int nol = no_layer - 1;
      if (nol < 0 || nol >= n_lay) { annp_b200_free_potential(out); set_err(err, errlen, "layer index out of range in weight block"); return ANNP_B200_EIO; }
      if (!flag_wb) {
        for (int i = 0; i < nrow_w; i++) {
          if (!std::getline(fin, t)) { annp_b200_free_potential(out); set_err(err, errlen, "potential file truncated in weights"); return ANNP_B200_EIO; }
          std::vector<double> v = row_values(t, true);
          if ((int) v.size() > n_sf || (int) v.size() > (ncol_w > n_sf ? ncol_w : n_sf)) { annp_b200_free_potential(out); set_err(err, errlen, "weight row longer than nsf"); return ANNP_B200_EIO; }
          double *dst = out->weight_all + (((size_t) type_elem * n_lay + nol) * n_nod + i) * n_sf;
          for (size_t k = 0; k < v.size(); k++) dst[k] = v[k];
        }
      } else {
        for (int i = 0; i < nrow_b; i++) {
          if (!std::getline(fin, t)) { annp_b200_free_potential(out); set_err(err, errlen, "potential file truncated in bias"); return ANNP_B200_EIO; }
          std::vector<double> v = row_values(t, true);
          if ((int) v.size() > n_nod) { annp_b200_free_potential(out); set_err(err, errlen, "bias row longer than nnod"); return ANNP_B200_EIO; }
          double *dst = out->bias_all + ((size_t) type_elem * n_lay + nol) * n_nod;
          for (size_t k = 0; k < v.size(); k++) dst[k] = v[k];
        }
      }
    }
    // the Ni copy stops the weight loop at "#coefficent of symmetry funciton" (first 5 characters compared,
    // ni/src/pair_annp.cpp:511-512) and then reads "#rad n" + npsf rows and "#angl n" + ntsf rows; numbers are
    // taken only after TAB+digit/'-' (the leading element names are skipped that way)
    if (t.compare(0, 5, "#coefficent of symmetry funciton", 0, 5) == 0) {
      out->has_sym_coeff = 1;
      if (!std::getline(fin, t)) break;
      for (int i = 0; i < out->npsf; i++) {
        if (!std::getline(fin, t)) { annp_b200_free_potential(out); set_err(err, errlen, "potential file truncated in radial coefficients"); return ANNP_B200_EIO; }
        int nv = 0;
        for (size_t j = 0; j < t.size(); j++)
          if (tab_number(t, j, true) && nv < 3) out->sym_coerad[i][nv++] = atof(t.c_str() + j + 1);
      }
      if (!std::getline(fin, t)) break;
      for (int i = 0; i < out->ntsf; i++) {
        if (!std::getline(fin, t)) { annp_b200_free_potential(out); set_err(err, errlen, "potential file truncated in angular coefficients"); return ANNP_B200_EIO; }
        int nv = 0;
        for (size_t j = 0; j < t.size(); j++)
          if (tab_number(t, j, true) && nv < 4) out->sym_coeang[i][nv++] = atof(t.c_str() + j + 1);
      }
      break;
    }
    if (fin.peek() == EOF) break;
  }
  return ANNP_B200_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// `.anna` files of the ANNA-ADP style: PairANNA_ADP::read_file (anna-gpu-lammps/bcc_fe/src/pair_anna_adp.cpp:392-637).
// Same tokenisation rules; the header is 19 + nelements lines: line 8+nel carries an extra Num_out field, there are no
// normalisation rows, line 14+nel holds "e_base<TAB>e_scal" (e_scal only after TAB+digit), 17+nel the number of global
// ADP parameters and 18+nel their values; the last weight block has nout rows.
extern "C" void anna_b200_free_potential(anna_b200_potential *pot) {
  if (!pot) return;
  free(pot->weight_all);
  free(pot->bias_all);
  pot->weight_all = nullptr;
  pot->bias_all = nullptr;
}

extern "C" int anna_b200_read_potential(const char *filename, int nelements_coeff, const char *const *elements_coeff,
                                        anna_b200_potential *out, char *err, int errlen) {
  (void) elements_coeff;
  if (!filename || !out) { set_err(err, errlen, "null argument"); return ANNP_B200_EINVAL; }
  memset(out, 0, sizeof(*out));
  std::ifstream fin(filename, std::ios::in);
  if (!fin.is_open()) {
    set_err(err, errlen, "Cannot open physically informed neural network potential file");   // pair_anna_adp.cpp:400
    return ANNP_B200_EIO;
  }
  std::string t;
  int nel = 0;
  for (int i = 0; i < 19 + nelements_coeff; i++) {
    if (!std::getline(fin, t)) { set_err(err, errlen, "potential file truncated in header"); return ANNP_B200_EIO; }
    if (i == 5) {
      nel = out->nelements = atoi(t.c_str());
      if (nel < 1 || nel > ANNP_B200_MAX_ELEMENTS) { set_err(err, errlen, "unsupported number of elements"); return ANNP_B200_EINVAL; }
    }
    if (i >= 6 && i < 6 + nel) {
      const int e = i - 6;
      out->id_elem[e] = atoi(t.c_str());
      std::string name;
      for (size_t j = 0; j < t.size(); j++) {
        if (isalpha((unsigned char) t[j])) name += t[j];
        if (tab_number(t, j, false)) out->mass[e] = atof(t.c_str() + j + 1);
      }
      snprintf(out->elements[e], sizeof(out->elements[e]), "%s", name.c_str());
    }
    if (i == 8 + nel) {
      out->ntl = atoi(t.c_str());
      int np = 1;
      for (size_t j = 0; j < t.size(); j++) {
        if (tab_number(t, j, false)) {
          const char *p = t.c_str() + j + 1;
          if (np == 1) out->nhl = atoi(p);
          if (np == 2) out->nnod = atoi(p);
          if (np == 3) out->nout = atoi(p);
          if (np == 4) out->nsf = atoi(p);
          if (np == 5) out->npsf = atoi(p);
          if (np == 6) out->ntsf = atoi(p);
          if (np == 7) out->cut = atof(p);
          np++;
        }
      }
      if (out->ntl < 2 || out->ntl - 1 > ANNP_B200_MAX_LAYERS || out->nnod < 1 || out->nnod > ANNP_B200_MAX_NOD ||
          out->nsf < 1 || out->nsf > ANNP_B200_MAX_SF || out->nout < 1 || out->nout > out->nnod) {
        set_err(err, errlen, "network dimensions outside the supported range");
        return ANNP_B200_EINVAL;
      }
    }
    if (i == 11 + nel) {
      int nact = 0;
      const int nlayer = out->ntl - 1;
      for (size_t j = 0; j < t.size(); j++) {
        const char a = t[j], b = (j + 1 < t.size()) ? t[j + 1] : '\0';
        if (a == 'C' && b == 'h') out->flagsym = 0;
        if ((a == 'B' && b == 'e') || (a == 'B' && b == 'P')) out->flagsym = 1;
        if (a == 'C' && b == 'u') out->flagsym = 2;
        int act = -1;
        if (a == 'l' && b == 'i') act = 0;
        if (a == 'h' && b == 'y') act = 1;
        if (a == 's' && b == 'i') act = 2;
        if (a == 'm' && b == 'o') act = 3;
        if (a == 't' && b == 'a') act = 4;
        if (act >= 0) {
          if (nact >= nlayer || nact >= ANNP_B200_MAX_LAYERS) { set_err(err, errlen, "more activation keywords than layers"); return ANNP_B200_EINVAL; }
          out->flagact[nact++] = act;
        }
      }
    }
    if (i == 14 + nel) {
      out->e_base = atof(t.c_str());
      for (size_t j = 0; j < t.size(); j++)
        if (tab_number(t, j, false)) out->e_scal = atof(t.c_str() + j + 1);
    }
    if (i == 17 + nel) {
      out->ngp = atoi(t.c_str());
      if (out->ngp < 1 || out->ngp > ANNA_B200_MAX_GPARAMS) { set_err(err, errlen, "unsupported number of ADP parameters"); return ANNP_B200_EINVAL; }
    }
    if (i == 18 + nel) {
      std::vector<double> v = row_values(t, true);
      if ((int) v.size() > ANNA_B200_MAX_GPARAMS) { set_err(err, errlen, "too many ADP parameters"); return ANNP_B200_EINVAL; }
      for (size_t k = 0; k < v.size(); k++) out->gparams[k] = v[k];
    }
  }
  if (out->ntl < 2) { set_err(err, errlen, "potential file has no network record"); return ANNP_B200_EIO; }

  const int n_lay = out->ntl - 1, n_nod = out->nnod, n_sf = out->nsf;
  out->weight_all = (double *) calloc((size_t) nel * n_lay * n_nod * n_sf, sizeof(double));
  out->bias_all = (double *) calloc((size_t) nel * n_lay * n_nod, sizeof(double));
  if (!out->weight_all || !out->bias_all) { anna_b200_free_potential(out); set_err(err, errlen, "out of host memory"); return ANNP_B200_ENOMEM; }
  while (true) {
    if (!std::getline(fin, t)) break;
    if (t.size() >= 2 && t[0] == '#' && isdigit((unsigned char) t[1])) {
      int no_layer = 0;
      bool flag_wb = false;
      for (size_t i = 0; i < t.size(); i++) {
        if (t[i] > 47 && t[i] < 58) { no_layer *= 10; no_layer += t[i] - 48; }
        if (t[i] == 'w') flag_wb = false;
        if (t[i] == 'b') flag_wb = true;
      }
      int nrow_w = n_nod;
      if (no_layer == out->ntl - 1) nrow_w = out->nout;
      const int nol = no_layer - 1;
      if (nol < 0 || nol >= n_lay) { anna_b200_free_potential(out); set_err(err, errlen, "layer index out of range in weight block"); return ANNP_B200_EIO; }
      if (!flag_wb) {
        for (int i = 0; i < nrow_w; i++) {
          if (!std::getline(fin, t)) { anna_b200_free_potential(out); set_err(err, errlen, "potential file truncated in weights"); return ANNP_B200_EIO; }
          std::vector<double> v = row_values(t, true);
          if ((int) v.size() > n_sf) { anna_b200_free_potential(out); set_err(err, errlen, "weight row longer than nsf"); return ANNP_B200_EIO; }
          double *dst = out->weight_all + (((size_t) nol) * n_nod + i) * n_sf;
          for (size_t k = 0; k < v.size(); k++) dst[k] = v[k];
        }
      } else {
        if (!std::getline(fin, t)) { anna_b200_free_potential(out); set_err(err, errlen, "potential file truncated in bias"); return ANNP_B200_EIO; }
        std::vector<double> v = row_values(t, true);
        if ((int) v.size() > n_nod) { anna_b200_free_potential(out); set_err(err, errlen, "bias row longer than nnod"); return ANNP_B200_EIO; }
        double *dst = out->bias_all + ((size_t) nol) * n_nod;
        for (size_t k = 0; k < v.size(); k++) dst[k] = v[k];
      }
    }
    if (fin.peek() == EOF) break;
  }
  return ANNP_B200_OK;
}
