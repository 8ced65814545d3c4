useOpenMP = False
gitSHA = ""
mask_size = 256
