useOpenMP = False
gitSHA = ""
