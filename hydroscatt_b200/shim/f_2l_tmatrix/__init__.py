"""Drop-in for hydroscatt's f2py package ``f_2l_tmatrix`` (reference README.md:11-19)."""
