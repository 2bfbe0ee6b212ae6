"""String constants of the public API.  The names and values are the reference's (matrix/constants.py:4-35): they are
what callers pass as `return_type=` / `permutation=` and what the save format writes."""

# decomposition types
BASE_DECOMPOSITION_TYPE = 'base'
DECOMPOSITION_TYPES = ('LDL', 'LDL_compressed', 'LL')
""" Supported types of decompositions. """
LDL_DECOMPOSITION_TYPE, LDL_DECOMPOSITION_COMPRESSED_TYPE, LL_DECOMPOSITION_TYPE = DECOMPOSITION_TYPES

# static permutation methods: by the diagonal values or their absolute values, decreasing or increasing
NO_PERMUTATION_METHOD = 'none'
DIAGONAL_VALUES_PERMUTATION_METHODS = tuple(
    '{}_{}diagonal_values'.format(direction, what) for what in ('', 'absolute_') for direction in ('decreasing', 'increasing'))
(DECREASING_DIAGONAL_VALUES_PERMUTATION_METHOD, INCREASING_DIAGONAL_VALUES_PERMUTATION_METHOD,
 DECREASING_ABSOLUTE_DIAGONAL_VALUES_PERMUTATION_METHOD,
 INCREASING_ABSOLUTE_DIAGONAL_VALUES_PERMUTATION_METHOD) = DIAGONAL_VALUES_PERMUTATION_METHODS
UNIVERSAL_PERMUTATION_METHODS = (NO_PERMUTATION_METHOD,) + DIAGONAL_VALUES_PERMUTATION_METHODS
""" Supported permutation methods for dense matrices. """
SPARSE_ONLY_PERMUTATION_METHODS = ()
""" The sparse (CHOLMOD) backend is out of scope of this package (SURVEY.md section 2). """

# save / load: a tar file `<name>.dec` holding `type.txt` and one `attribute_<name>.dense.npz` per attribute
DECOMPOSITION_FILENAME_EXTENSION = 'dec'
DECOMPOSITION_TYPE_FILENAME = 'type.txt'
DECOMPOSITION_ATTRIBUTE_FILENAME = 'attribute_{attribute_name}.{file_extension}'
DECOMPOSITION_ATTRIBUTE_DENSE_FILE_EXTENSION, DECOMPOSITION_ATTRIBUTE_SPARSE_FILE_EXTENSION = 'dense.npz', 'sparse.npz'
