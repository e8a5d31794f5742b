from mutation_simulator_b200.__main__ import main

if __name__ == "__main__":
    main()
