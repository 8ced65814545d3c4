initialize = True
threads = True
thread_level = 'multiple'
finalize = None
