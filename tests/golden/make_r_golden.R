#!/usr/bin/env Rscript
# Pins the CPU oracle to the REAL reference, wherever R is available (it is not in the build image nor on the GPU box:
# SURVEY.md F4, 8c last row).  Runs the reference's own TADpole() -- the installed package, or its sources under
# $TADPOLE_REFERENCE (default /root/reference) -- on a matrix file written by tests/golden/r_golden.py and dumps what the
# oracle is compared with:
#     Rscript tests/golden/make_r_golden.R <matrix.tsv> <out.json> [max_pcs] [centromere_search]
# Needs rioja, fpc, bigmemory, Matrix, foreach, doParallel (the reference's Imports) and jsonlite.  Never run so far.
args <- commandArgs(trailingOnly = TRUE)
stopifnot(length(args) >= 2)
mat_file <- args[1]; out_file <- args[2]
max_pcs <- if (length(args) >= 3) as.integer(args[3]) else 200L
cen <- length(args) >= 4 && as.logical(args[4])
suppressPackageStartupMessages(library(jsonlite))
if (requireNamespace("TADpole", quietly = TRUE)) {
    TADpole <- TADpole::TADpole
} else {
    ref <- Sys.getenv("TADPOLE_REFERENCE", "/root/reference")
    for (f in c("TADpole.R", "DiffT.R")) source(file.path(ref, "R", f))
}
pdf(NULL)                                  # load_mat draws two figures (R/TADpole.R:24-53)
n <- length(strsplit(readLines(mat_file, n = 1), "\t")[[1]])
tp <- TADpole(mat_file, max_pcs = max_pcs, min_clusters = 2, bad_frac = 0.01, chr = "chrG", start = 1, end = n,
              resol = 1, centromere_search = cen)
dendro <- function(d) list(merge = d$merge, height = d$height, seqdist = d$seqdist, labels = d$labels)
tabs <- function(cl) lapply(cl, function(df) unname(as.matrix(df[, c("start", "end")])))
part <- function(x, nm) list(n_pcs = x$n_pcs, optimal_n_clusters = x$optimal_n_clusters, dendro = dendro(x$dendro),
                             clusters = tabs(x[[nm]]))
out <- if (cen) list(p = part(tp$p, "cluster"), q = part(tp$q, "cluster"),
                     merging_arms = unname(as.matrix(tp$merging_arms[, c("start", "end")])))
       else c(part(tp, "clusters"), list(scores = tp$scores))
out$versions <- list(R = R.version.string, rioja = as.character(packageVersion("rioja")),
                     fpc = as.character(packageVersion("fpc")))
writeLines(toJSON(out, digits = NA, na = "null", auto_unbox = TRUE), out_file)
