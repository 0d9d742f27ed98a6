# What Rcpp::compileAttributes() generates from rpkg/src/InterfaceUtils.cpp (run it after copying that file into
# src/; it also regenerates src/RcppExports.cpp with the matching '_DepecheR_*' entry points and their registration).
# The first four functions are byte-for-byte the reference's R/RcppExports.R:4-18.

sparse_k_means <- function(X, k, reg, no_zero, seed_off_set = 0L) {
    .Call('_DepecheR_sparse_k_means', PACKAGE = 'DepecheR', X, k, reg, no_zero, seed_off_set)
}

grid_search <- function(X, k, reg, iterations, bootstrapSamples, seed_off_set = 0L) {
    .Call('_DepecheR_grid_search', PACKAGE = 'DepecheR', X, k, reg, iterations, bootstrapSamples, seed_off_set)
}

allocate_points <- function(X, mu, no_zero) {
    .Call('_DepecheR_allocate_points', PACKAGE = 'DepecheR', X, mu, no_zero)
}

rand_index <- function(inds1, inds2, k) {
    .Call('_DepecheR_rand_index', PACKAGE = 'DepecheR', inds1, inds2, k)
}

set_engine_seed <- function(seed) {
    invisible(.Call('_DepecheR_set_engine_seed', PACKAGE = 'DepecheR', seed))
}

dataset_create <- function(X) {
    .Call('_DepecheR_dataset_create', PACKAGE = 'DepecheR', X)
}

dataset_destroy <- function(handle) {
    invisible(.Call('_DepecheR_dataset_destroy', PACKAGE = 'DepecheR', handle))
}

dataset_create_scaled <- function(X, log2Off, centerMode, centers, scaleMode, scaleValue) {
    .Call('_DepecheR_dataset_create_scaled', PACKAGE = 'DepecheR', X, log2Off, centerMode, centers, scaleMode, scaleValue)
}

dataset_allocate_points_ranked <- function(handle, mu, no_zero) {
    .Call('_DepecheR_dataset_allocate_points_ranked', PACKAGE = 'DepecheR', handle, mu, no_zero)
}

dataset_gather <- function(handle, rows) {
    .Call('_DepecheR_dataset_gather', PACKAGE = 'DepecheR', handle, rows)
}

dataset_allocate_points <- function(handle, mu, no_zero) {
    .Call('_DepecheR_dataset_allocate_points', PACKAGE = 'DepecheR', handle, mu, no_zero)
}

grid_search_each <- function(handle, k, reg, iterations, bootstrapSamples, seed_off_set = 0L) {
    .Call('_DepecheR_grid_search_each', PACKAGE = 'DepecheR', handle, k, reg, iterations, bootstrapSamples, seed_off_set)
}

penalty_opt <- function(handle, penalties, k, sampleSize, nCores, maxIter, minARIImprovement, optimARI, seed_off_set = 0L) {
    .Call('_DepecheR_penalty_opt', PACKAGE = 'DepecheR', handle, penalties, k, sampleSize, nCores, maxIter, minARIImprovement, optimARI, seed_off_set)
}

select_solution <- function(handle, allSolutions, k) {
    .Call('_DepecheR_select_solution', PACKAGE = 'DepecheR', handle, allSolutions, k)
}
