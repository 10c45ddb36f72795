// Command line of the B200 build: the reference's options (main.cpp:52-154) in front of
// Assembler::runBayesemblerBatched (assembler.hpp), from the splice-graph stage on.
//
//     bayesembler_b200_cli --graph-file graphs.txt -o out/ --seed 1 --frag-mean 200 --frag-sd 20 --output-mode full
#include <sys/stat.h>

#include "assembler.hpp"

using namespace bayesembler_b200;

static void make_directories(const std::string &dir) {  // boost::filesystem::create_directories (main.cpp:131-141)
    for (size_t i = 1; i <= dir.size(); i++)
        if (i == dir.size() || dir[i] == '/') mkdir(dir.substr(0, i).c_str(), 0777);
}

int main(int argc, char *const argv[]) {
    OptionsContainer options_variables;
    std::cout << "\nYou are using the B200 build of the Bayesembler v1.2.0 sampler path (splice-graph stage on).\n" << std::endl;
    const int rc = parseCommandLine(argc, argv, &options_variables, std::cout, std::cerr);
    if (rc != 0) return 1;
    try {
        const size_t slash = options_variables.output_prefix.rfind('/');
        if (slash != std::string::npos && slash > 0) {
            std::cout << "Creating output directory " << options_variables.output_prefix.substr(0, slash) << std::endl;
            make_directories(options_variables.output_prefix.substr(0, slash));
        }
        if (!options_variables.estimate_fragment_length)
            std::cout << "Using fragment length mean " << options_variables.frag_mean << " and sd " << options_variables.frag_sd << std::endl;
        Assembler assembler(options_variables);
        assembler.runBayesemblerBatched(std::cout);
    } catch (std::exception &e) {
        std::cerr << "ERROR: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
