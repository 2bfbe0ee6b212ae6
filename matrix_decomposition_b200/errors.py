"""Exception hierarchy of the public API.

Same class names, constructor arguments, attributes and messages as the reference's
matrix/errors.py:8-176, so that `except matrix.errors.X` clauses keep working.
"""


class BaseError(Exception):
    """ Base exception of this package (errors.py:8-16). """

    def __init__(self, message):
        self.message = message
        super().__init__(message)

    def __str__(self):
        return self.message


# *** matrix exceptions (errors.py:21-76) *** #

class MatrixError(BaseError):
    def __init__(self, matrix, message=None):
        self.matrix = matrix
        if message is None:
            message = 'Error with matrix with shape {}'.format(self.matrix.shape)
        super().__init__(message)


class MatrixNotSquareError(MatrixError):
    def __init__(self, matrix):
        super().__init__(matrix, message='Matrix with shape {} is not a square matrix.'.format(matrix.shape))


class MatrixNotFiniteError(MatrixError):
    def __init__(self, matrix):
        super().__init__(matrix, message='Matrix with shape {} has not finite entries.'.format(matrix.shape))


class MatrixSingularError(MatrixError):
    def __init__(self, matrix):
        super().__init__(matrix, message='Matrix with shape {} is singular.'.format(matrix.shape))


class MatrixNotHermitianError(MatrixError):
    def __init__(self, matrix, i=None, j=None):
        message = 'Matrix with shape {} is not Hermitian.'.format(matrix.shape)
        if i is not None and j is not None:
            if i != j:
                message += (' Its value at index ({i}, {j}) is {A_i_j} and its value at index ({j}, {i}) is {A_j_i}.'
                            ).format(i=i, j=j, A_i_j=matrix[i, j], A_j_i=matrix[j, i])
            else:
                message += ' Its {i}-th diagonal value {A_i_i} is complex.'.format(i=i, A_i_i=matrix[i, i])
        super().__init__(matrix, message=message)


class MatrixComplexDiagonalValueError(MatrixNotHermitianError):
    def __init__(self, matrix, i=None):
        super().__init__(matrix, i=i, j=i)


# *** decomposition exceptions (errors.py:81-122) *** #

class DecompositionError(BaseError):
    def __init__(self, decomposition, message=None):
        self.decomposition = decomposition
        if message is None:
            message = 'Error with decomposition {}'.format(self.decomposition)
        super().__init__(message)


class DecompositionNotFiniteError(DecompositionError):
    def __init__(self, decomposition):
        super().__init__(decomposition, 'Decomposition {} has non-finite entries.'.format(decomposition))


class DecompositionSingularError(DecompositionError):
    def __init__(self, decomposition):
        super().__init__(decomposition, 'Decomposition {} represents a singular matrix.'.format(decomposition))


class DecompositionInvalidFile(DecompositionError, OSError):
    def __init__(self, filename):
        self.filename = filename
        DecompositionError.__init__(self, None, 'File {} is not a valid decomposition file.'.format(filename))


class DecompositionInvalidDecompositionTypeFile(DecompositionInvalidFile):
    def __init__(self, filename, type_file, type_needed):
        self.filename = filename
        DecompositionError.__init__(
            self, None, 'File {} contains a decomposition of type {} but type {} is needed.'.format(
                filename, type_file, type_needed))


# *** decomposition not calculable (errors.py:127-157) *** #

class NoDecompositionPossibleError(BaseError):
    def __init__(self, base, desired_type):
        self.desired_type = desired_type
        self.base = base
        super().__init__('Base {} can not be converted to type {}.'.format(base, desired_type))


class NoDecompositionPossibleWithProblematicSubdecompositionError(NoDecompositionPossibleError):
    def __init__(self, base, desired_type, problematic_leading_principal_submatrix_index, subdecomposition=None):
        super().__init__(base, desired_type)
        self.problematic_leading_principal_submatrix_index = problematic_leading_principal_submatrix_index
        if subdecomposition is not None:
            self.subdecomposition = subdecomposition


class NoDecompositionPossibleTooManyEntriesError(NoDecompositionPossibleError):
    def __init__(self, matrix, desired_type):
        super().__init__(matrix, desired_type)


class NoDecompositionConversionImplementedError(NoDecompositionPossibleError):
    def __init__(self, decomposition, desired_type):
        super().__init__(decomposition, desired_type)


# *** this package only *** #

class NativeLibraryError(BaseError):
    """ libmdb200.so is missing or a CUDA call failed.  There is no CPU fallback. """
