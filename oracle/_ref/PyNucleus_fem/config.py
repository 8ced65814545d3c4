useOpenMP = False
gitSHA = ""
use_metis = True
