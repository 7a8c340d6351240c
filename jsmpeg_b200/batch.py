def bind_batch_abi(lib):
    return lib
