def assemble_matrix(form, qr_data):
    raise NotImplementedError
assemble_vector = assemble_scalar = assemble_matrix
