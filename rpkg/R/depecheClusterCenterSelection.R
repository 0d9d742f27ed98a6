# Replaces R/depecheClusterCenterSelection.R of the reference (same arguments and return value).
# The reference allocates the selection set to every stored solution with one dAllocate call each (:11-16) and then starts a
# SOCK cluster for the M x M rand_index calls (:22-30).  Here the selection set is uploaded once and select_solution
# (dpr_dataset_select_solution) allocates it to all M solutions in one pass over the cells and scores all pairs on the device.
depecheClusterCenterSelection <- function(allSolutions, selectionDataSet, k,
                                          nCores) {
    handle <- dataset_create(data.matrix(selectionDataSet, rownames.force = NA))
    on.exit(dataset_destroy(handle), add = TRUE)
    sel <- select_solution(handle, lapply(allSolutions, as.matrix), k)
    optimalClusterCenters <- as.matrix(allSolutions[[sel$best]])
    colnames(optimalClusterCenters) <- colnames(selectionDataSet)
    # Remove all rows and columns that do not contain any information (:41-45)
    keepRows <- which(rowSums(optimalClusterCenters) != 0)
    keepCols <- which(colSums(optimalClusterCenters) != 0)
    optimalClusterCenters[keepRows, keepCols, drop = FALSE]
}
