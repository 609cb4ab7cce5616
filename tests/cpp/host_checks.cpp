// host_checks.cpp -- CPU-only checks of the C++ host layer (no GPU needed):
//  * b200::RowLayout / sigmoid_coeffs print their packing so that the Python side can compare (tests/test_shim.py)
//  * creating a context without a CUDA device fails loudly (no CPU fallback), with the engine's message
#include <iostream>

#include "ckks_b200_lr.h"

int main(int argc, char **argv) {
    using namespace std;
    if (argc > 1 && string(argv[1]) == "layout") {
        const int R = 5, C = 3;
        vector<vector<double>> X(R, vector<double>(C));
        for (int i = 0; i < R; i++)
            for (int j = 0; j < C; j++) X[i][j] = 10 * i + j + 1;
        b200::RowLayout lay(R, C, 16);
        cout.precision(17);
        for (int i = 0; i < R; i++) {
            for (double v : lay.row(X, i)) cout << v << " ";
            cout << "\n";
        }
        for (int j = 0; j < C; j++) {
            for (double v : lay.column(X, j)) cout << v << " ";
            cout << "\n";
        }
        for (double v : lay.weights({0.5, -1.5, 2.5})) cout << v << " ";
        cout << "\n";
        for (double v : lay.labels({1, 0, 1, 1, 0})) cout << v << " ";
        cout << "\n";
        for (int d : {3, 5, 7}) {
            for (double v : b200::sigmoid_coeffs(d)) cout << v << " ";
            cout << "\n";
        }
        return 0;
    }
    seal::EncryptionParameters params(seal::scheme_type::CKKS);
    params.set_poly_modulus_degree(8192);
    params.set_coeff_modulus(seal::CoeffModulus::Create(8192, {60, 40, 40, 60}));
    try {
        auto context = seal::SEALContext::Create(params);
        seal::KeyGenerator keygen(context);
        cout << "context created" << endl;
        return 0;
    } catch (const std::exception &e) {
        cout << "exception: " << e.what() << endl;
        return 3;
    }
}
