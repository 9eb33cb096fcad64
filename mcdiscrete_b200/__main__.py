from .doScatter import main

main()
