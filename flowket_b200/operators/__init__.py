from .operator import Operator, OperatorOnGrid, cube_shape
from .heisenberg import Heisenberg
from .ising import Ising
from .j1j2 import J1J2, j1j2_two_dim_operator
from .flattened import FlattenedOperator
