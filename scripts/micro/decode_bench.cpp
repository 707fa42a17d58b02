// Decode-only benchmark of the front end (no GPU): records/s of AlignmentDecoder::decode on a SAM / BAM file.
//   g++ -O2 -std=c++17 -pthread -I slimm_b200/csrc/frontend scripts/micro/decode_bench.cpp -lz -o /tmp/decode_bench
//   /tmp/decode_bench file.sam [threads] [grouped=1]
#include "alignment_decoder.hpp"
#include <chrono>
#include <iostream>
using namespace slimm_fe;
int main(int argc, char **argv)
{
    if (argc < 2) return 1;
    const int threads = argc > 2 ? atoi(argv[2]) : (int)std::thread::hardware_concurrency();
    const bool grouped = argc > 3 ? atoi(argv[3]) != 0 : true;
    for (int rep = 0; rep < 3; ++rep) {
        AlignmentDecoder dec;
        std::string err;
        auto t0 = std::chrono::steady_clock::now();
        if (!dec.open(argv[1], err)) { std::cerr << err << "\n"; return 1; }
        auto t1 = std::chrono::steady_clock::now();
        const size_t cap = 1u << 20;
        std::vector<uint32_t> rid(cap), ref(cap);
        std::vector<int32_t> pos(cap);
        RecordBatch batch{rid.data(), ref.data(), pos.data(), 0, cap};
        DecodeStats st;
        uint64_t sum = 0;
        const bool ok = dec.decode(threads, batch, [&](RecordBatch b) { sum += b.n ? b.read_id[b.n - 1] : 0; b.n = 0; return b; }, st, err, grouped);
        auto t2 = std::chrono::steady_clock::now();
        const double open_s = std::chrono::duration<double>(t1 - t0).count(), dec_s = std::chrono::duration<double>(t2 - t1).count();
        std::cout << "rep " << rep << ": ok=" << ok << " records " << st.records_kept << " reads " << st.reads << " open " << open_s << " s decode " << dec_s << " s = "
                  << st.records_kept / dec_s / 1e6 << " M records/s (" << threads << " threads, " << (grouped ? "grouped" : "exact") << ") " << err << " [" << sum << "]\n";
    }
}
