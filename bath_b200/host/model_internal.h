// model_internal.h -- the query model object shared by the host-side sources.
#pragma once
#include "host_internal.h"

struct bathhost_model {
  bathhost::CoreModel      hmm;
  bathhost::NullModel      bg;
  int                      ct = 1;
  int                      maxl_in_file = -1;
  bathhost::FsProfile      gm3, gm5;
  bathhost::FsOddsProfile  om3, om5;
  bathhost::ProteinProfile prot;
};
