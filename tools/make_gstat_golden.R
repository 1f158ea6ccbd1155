#!/usr/bin/env Rscript
# Pins the kriging oracle (oracle/twx_oracle.py: ked_gstat, gcdist_sp) and the CUDA kernel to REAL gstat.
#
# Needs R with the packages the reference uses (sp, gstat; INSTALL.rst:56,59 pins gstat 1.0-25 / sp 1.1-1) and a checkout
# of jaredwo/topowx.  It sources the reference's own twx/interp/rpy/interp.R, sets FORMULA exactly as
# twx/interp/interp_tair.py:83 does, and calls krig_meantair (interp.R:198-270) on the committed neighbourhoods
# tests/golden/gstat_krige_inputs_{nghs,pts}.csv.  Output: tests/golden/gstat_krige.csv (case, mean, var, gstat version);
# commit it and tests/test_oracle_vs_gstat.py stops skipping.
#
#   Rscript tools/make_gstat_golden.R /path/to/topowx [repo root]
args <- commandArgs(trailingOnly = TRUE)
ref <- ifelse(length(args) >= 1, args[1], "/root/reference")
root <- ifelse(length(args) >= 2, args[2], ".")
source(file.path(ref, "twx", "interp", "rpy", "interp.R"))
FORMULA <- build_formula("tair", c("longitude", "latitude", "elevation", "lst"))   # KRIG_TREND_VARS, interp_tair.py:46,83
nghs <- read.csv(file.path(root, "tests", "golden", "gstat_krige_inputs_nghs.csv"))
pts <- read.csv(file.path(root, "tests", "golden", "gstat_krige_inputs_pts.csv"))
out <- data.frame(case = pts$case, mean = NA_real_, var = NA_real_)
for (i in seq_len(nrow(pts))) {
  g <- nghs[nghs$case == pts$case[i], ]
  pt <- c(pts$longitude[i], pts$latitude[i], pts$elevation[i], pts$tdi[i], pts$lst[i])
  r <- tryCatch(krig_meantair(g$longitude, g$latitude, g$elevation, g$tdi, g$lst, g$tair, g$ngh_wgt, pt,
                              pts$nug[i], pts$psill[i], pts$range[i]),
                error = function(e) c(NA, NA, 1))
  out$mean[i] <- r[1]
  out$var[i] <- r[2]
}
out$gstat <- as.character(packageVersion("gstat"))
out$sp <- as.character(packageVersion("sp"))
write.csv(format(out, digits = 17), file.path(root, "tests", "golden", "gstat_krige.csv"), row.names = FALSE, quote = FALSE)
cat(sprintf("wrote %d cases (%d failed) with gstat %s\n", nrow(out), sum(is.na(out$mean)), out$gstat[1]))
