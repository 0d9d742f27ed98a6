# One-process driver of the penalty search for the B200 build: replaces R/depechePenaltyOpt.R of the reference
# (same arguments, same returned list: bestPenalty, meanOptimDf(ARI, nClust; row names = rounded penalties),
# bestClusterCenters).
#
# The reference starts nCores SOCK workers and, in every set of its while loop, ships the whole matrix to each of them for a
# grid_search call (reference R/depechePenaltyOpt.R:37-38,50-55).  Ten R processes would be ten CUDA contexts and ten uploads per
# set.  Here the matrix is uploaded ONCE (dataset_create -> external pointer into HBM), and the whole adaptive loop — the sets,
# the zeroing of trivial solutions (:75-78), mean / sd per penalty (:89-100), the pruning of penalties (:123-140), the stdChange
# stopping rule (:145-156, :47-48) and the choice of the best penalty (:184-192) — runs inside one call (penalty_opt ->
# dpr_dataset_penalty_opt), each set as one batch of nCores x length(usedPenalties) problems on the GPU.  `nCores` keeps its
# meaning as the number of worker results per set.
depechePenaltyOpt <- function(inDataFrameUsed, k, maxIter, minARIImprovement,
                              sampleSize, penalties, createOutput,
                              disableWarnings = FALSE, optimARI, nCores,
                              plotDir) {
    roundPenalties <- round(penalties, digits = 1)
    dataMat <- data.matrix(inDataFrameUsed, rownames.force = NA)

    handle <- dataset_create(dataMat)
    on.exit(dataset_destroy(handle), add = TRUE)
    ptm <- proc.time()
    res <- penalty_opt(handle, as.numeric(penalties), k, sampleSize, nCores,
                       maxIter, minARIImprovement, optimARI, 0L)
    fullTime <- proc.time() - ptm
    message("The optimization was iterated ", res$workerIterations,
            " times in ", round(fullTime[3]), " seconds.")

    if (!disableWarnings && bitwAnd(res$warn, 1L) != 0) {
        warning("The maximum number of iterations was reached before stable
                optimal solution was found")
    }
    bestPenalty <- res$bestPenalty
    if (length(penalties) > 1 && !disableWarnings) {
        if (bitwAnd(res$warn, 2L) != 0) {
            warning("The lowest penalty was the most optimal in the range. ",
                    "This might either be due a suboptimal set of penalties,",
                    " or to the use of a too small sample size")
        }
        if (bitwAnd(res$warn, 4L) != 0) {
            warning("The highest penalty was the most optimal in the range. ",
                    "This might either be due a suboptimal set of penalties,",
                    " or to the use of a too small sample size")
        }
    }
    meanARIVec <- res$ARI

    # The diagnostic plot of the reference (:211-238), unchanged
    if (createOutput) {
        pdf(file.path(plotDir, "ARI_as_a_function_of_penalties.pdf"))
        par(mar = c(5, 4, 4, 6) + 0.1)
        plot(log10(roundPenalties), meanARIVec,
            pch = 16, axes = FALSE,
            ylim = c(0, 1), xlab = "", ylab = "", type = "b", col = "black",
            main = "Difference between samples as a function of penalties"
        )
        axis(2, ylim = c(0, 1), col = "black", las = 1)
        mtext("Adjusted rand index (ARI) between data resamplings",
            side = 2, line = 2.5
        )
        graphics::box()
        axis(1, pretty(range(log10(roundPenalties)), n = 10))
        mtext("Log10 of penalty values", side = 1, col = "black", line = 2.5)
        legend("topleft",
            legend = "ARI (high is good)", text.col = "black",
            pch = c(16, 15), col = "black"
        )
        dev.off()
    }

    meanOptimDf <- data.frame("ARI" = meanARIVec, "nClust" = res$nClust)
    row.names(meanOptimDf) <- roundPenalties
    bestClusterCenters <- lapply(res$bestClusterCenters, function(x) {
        colnames(x) <- colnames(dataMat)
        x
    })
    list("bestPenalty" = bestPenalty, meanOptimDf, bestClusterCenters)
}

# The legacy name (reference R/dOptPenalty.R is the same function, used by tests/testthat/test_dOptPenalty.R:6)
dOptPenalty <- depechePenaltyOpt
