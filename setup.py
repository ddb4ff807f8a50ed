from setuptools import find_packages, setup

setup(
    name='arrow_matrix_b200',
    version='0.1.0',
    packages=find_packages(include=['arrow_matrix_b200', 'arrow_matrix_b200.*']),
    package_data={'arrow_matrix_b200': ['libarrow_b200.so', 'csrc/*.cu']},
    entry_points={'console_scripts': ['spmm_arrow=arrow_matrix_b200.cli:main',
                                        'arrow_decompose=arrow_matrix_b200.decompose_cli:main',
                                        'spmm_petsc=arrow_matrix_b200.baseline.petsc_cli:main',
                                        'spmm_15d=arrow_matrix_b200.baseline.spmm_15d_cli:main']},
)
