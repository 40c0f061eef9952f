// cvstorage.cpp -- the subset of cv::FileStorage the calibration ingest needs: "opencv-matrix" nodes and real scalars,
// in the YAML 1.0 and XML flavours OpenCV writes (read: src/mymulticalib.cpp:118-131, 182-196; write: :425-454).
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>

#include "host_impl.hpp"

namespace mccba {

namespace {
// dt = [channels]type, e.g. "d", "f", "2f", "3d"
void parse_dt(const std::string& dt, int& channels, char& depth)
{
    channels = 1;
    depth = 'd';
    size_t i = 0;
    while (i < dt.size() && std::isspace((unsigned char)dt[i])) ++i;
    int ch = 0;
    while (i < dt.size() && std::isdigit((unsigned char)dt[i])) { ch = ch * 10 + (dt[i] - '0'); ++i; }
    if (ch > 0) channels = ch;
    if (i < dt.size()) depth = dt[i];
}
double parse_real(const std::string& tok)
{
    // YAML spellings of the special values (".Nan", ".Inf", "-.Inf")
    if (tok.find(".Nan") != std::string::npos || tok.find(".nan") != std::string::npos) return std::nan("");
    if (tok.find(".Inf") != std::string::npos || tok.find(".inf") != std::string::npos) return tok[0] == '-' ? -HUGE_VAL : HUGE_VAL;
    return std::strtod(tok.c_str(), nullptr);
}
void parse_numbers(const std::string& s, std::vector<double>& out)
{
    std::string tok;
    for (char ch : s) {
        if (ch == ',' || std::isspace((unsigned char)ch) || ch == '[' || ch == ']') {
            if (!tok.empty()) { out.push_back(parse_real(tok)); tok.clear(); }
        } else tok.push_back(ch);
    }
    if (!tok.empty()) out.push_back(parse_real(tok));
}
std::string trim(const std::string& s)
{
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}
// YAML: position just behind "name:" when the key starts a line (top-level keys only)
size_t yaml_key(const std::string& t, const std::string& name)
{
    const std::string key = name + ":";
    size_t pos = 0;
    while ((pos = t.find(key, pos)) != std::string::npos) {
        if (pos == 0 || t[pos - 1] == '\n') return pos + key.size();
        pos += key.size();
    }
    return std::string::npos;
}
// value of "field: value" inside the node that starts at `from` (fields are indented lines)
bool yaml_field(const std::string& t, size_t from, size_t to, const std::string& field, std::string& value)
{
    const std::string key = field + ":";
    const size_t p = t.find(key, from);
    if (p == std::string::npos || p >= to) return false;
    size_t e = t.find('\n', p);
    if (e == std::string::npos) e = t.size();
    value = trim(t.substr(p + key.size(), e - p - key.size()));
    return true;
}
// XML: content between <name ...> and </name>; attrs receives the attribute text of the opening tag
bool xml_node(const std::string& t, const std::string& name, size_t from, size_t to, std::string& attrs, std::string& content)
{
    size_t pos = from;
    const std::string open = "<" + name;
    while ((pos = t.find(open, pos)) != std::string::npos && pos < to) {
        const char nx = pos + open.size() < t.size() ? t[pos + open.size()] : '\0';
        if (nx == '>' || nx == ' ' || nx == '\t' || nx == '\n' || nx == '/') break;
        pos += open.size();
    }
    if (pos == std::string::npos || pos >= to) return false;
    const size_t gt = t.find('>', pos);
    if (gt == std::string::npos) return false;
    attrs = t.substr(pos + open.size(), gt - pos - open.size());
    if (!attrs.empty() && attrs.back() == '/') { content.clear(); return true; }   // <name/>
    const std::string close = "</" + name + ">";
    const size_t end = t.find(close, gt);
    if (end == std::string::npos) return false;
    content = t.substr(gt + 1, end - gt - 1);
    return true;
}
}  // namespace

CvFileReader::CvFileReader(const std::string& path)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    text_ = ss.str();
    size_t i = 0;
    while (i < text_.size() && std::isspace((unsigned char)text_[i])) ++i;
    xml_ = i < text_.size() && text_[i] == '<';
}

bool CvFileReader::matrix(const std::string& name, CvMatrix& out) const
{
    out = CvMatrix();
    std::string srows, scols, sdt, sdata;
    if (xml_) {
        std::string attrs, body, a2;
        if (!xml_node(text_, name, 0, text_.size(), attrs, body)) return false;
        if (!xml_node(body, "rows", 0, body.size(), a2, srows) || !xml_node(body, "cols", 0, body.size(), a2, scols) ||
            !xml_node(body, "dt", 0, body.size(), a2, sdt) || !xml_node(body, "data", 0, body.size(), a2, sdata))
            return false;
    } else {
        const size_t from = yaml_key(text_, name);
        if (from == std::string::npos) return false;
        // the node ends at the next top-level key (a line that does not start with white space) or at the end
        size_t to = from;
        while (true) {
            const size_t nl = text_.find('\n', to);
            if (nl == std::string::npos) { to = text_.size(); break; }
            if (nl + 1 >= text_.size()) { to = text_.size(); break; }
            const char c = text_[nl + 1];
            if (!std::isspace((unsigned char)c) && c != ']') { to = nl + 1; break; }
            to = nl + 1;
        }
        if (!yaml_field(text_, from, to, "rows", srows) || !yaml_field(text_, from, to, "cols", scols) ||
            !yaml_field(text_, from, to, "dt", sdt))
            return false;
        const size_t dk = text_.find("data:", from);
        if (dk == std::string::npos || dk >= to) return false;
        const size_t lb = text_.find('[', dk), rb = text_.find(']', dk);
        if (lb == std::string::npos || rb == std::string::npos || rb < lb) return false;
        sdata = text_.substr(lb + 1, rb - lb - 1);
        if (!sdt.empty() && (sdt[0] == '"' || sdt[0] == '\'')) sdt = sdt.substr(1, sdt.size() - 2);
    }
    out.rows = std::atoi(trim(srows).c_str());
    out.cols = std::atoi(trim(scols).c_str());
    parse_dt(trim(sdt), out.channels, out.depth);
    parse_numbers(sdata, out.data);
    if ((long long)out.data.size() != (long long)out.rows * out.cols * out.channels) { out = CvMatrix(); return false; }
    return true;
}

bool CvFileReader::scalar(const std::string& name, double& out) const
{
    std::string v;
    if (xml_) {
        std::string attrs;
        if (!xml_node(text_, name, 0, text_.size(), attrs, v)) return false;
    } else {
        const size_t from = yaml_key(text_, name);
        if (from == std::string::npos) return false;
        size_t e = text_.find('\n', from);
        if (e == std::string::npos) e = text_.size();
        v = text_.substr(from, e - from);
    }
    v = trim(v);
    if (v.empty() || v[0] == '!' || v[0] == '<') return false;
    out = parse_real(v);
    return true;
}

std::string cv_format_real(double v, int digits)
{
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%.*g", digits, v);
    std::string s(buf);
    if (s.find_first_of(".eEni") == std::string::npos) s += ".";   // OpenCV writes "1." for integral reals
    return s;
}

void cv_write_mat_xml(std::ostream& os, const std::string& name, int rows, int cols, char dt, const double* data)
{
    os << "<" << name << " type_id=\"opencv-matrix\">\n  <rows>" << rows << "</rows>\n  <cols>" << cols << "</cols>\n  <dt>" << dt
       << "</dt>\n  <data>\n    ";
    size_t col = 4;
    const int digits = dt == 'f' ? 9 : 17;     // enough to round-trip CV_32F / CV_64F
    for (int i = 0; i < rows * cols; ++i) {
        const std::string s = cv_format_real(dt == 'f' ? (double)(float)data[i] : data[i], digits);
        if (i > 0) {
            if (col + 1 + s.size() > 72) { os << "\n    "; col = 4; }
            else { os << " "; ++col; }
        }
        os << s;
        col += s.size();
    }
    os << "</data></" << name << ">\n";
}

}  // namespace mccba
