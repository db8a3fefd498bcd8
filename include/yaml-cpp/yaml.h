// Shim for the part of yaml-cpp the planner configuration uses (yaml-cpp is absent from this image):
// YAML::LoadFile, Node::operator[] (key / index), Node::as<T>(), Node::size(), truthiness, and the
// exceptions AbstractPlanner.hpp catches (reference src/planners/include/abstract/AbstractPlanner.hpp:20-81,
// src/planners/src/wrappers/stomp/HandleStompConfig.cpp:7-63).  Supported YAML: nested block mappings,
// scalars, flow sequences ([a, b, c]), block sequences (- a), comments.  Enough for test/config/stomp.yml.
#pragma once
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace YAML {

class Exception : public std::runtime_error {
public:
    explicit Exception(const std::string& m) : std::runtime_error(m) {}
};
class ParserException : public Exception {
public:
    explicit ParserException(const std::string& m) : Exception(m) {}
};
class BadFile : public Exception {
public:
    explicit BadFile(const std::string& m) : Exception("bad file: " + m) {}
};
class BadConversion : public Exception {
public:
    explicit BadConversion(const std::string& m) : Exception("bad conversion: " + m) {}
};

class Node {
public:
    enum Kind { Undefined, Scalar, Sequence, Map };
    Node() : d_(std::make_shared<Data>()) {}

    bool IsDefined() const { return d_->kind != Undefined; }
    bool IsScalar() const { return d_->kind == Scalar; }
    bool IsSequence() const { return d_->kind == Sequence; }
    bool IsMap() const { return d_->kind == Map; }
    explicit operator bool() const { return IsDefined(); }
    bool operator!() const { return !IsDefined(); }
    size_t size() const { return d_->kind == Sequence ? d_->seq.size() : d_->kind == Map ? d_->map.size() : 0; }

    const Node operator[](const std::string& key) const
    {
        if (d_->kind != Map) return Node();
        auto it = d_->map.find(key);
        return it == d_->map.end() ? Node() : it->second;
    }
    const Node operator[](const char* key) const { return (*this)[std::string(key)]; }
    const Node operator[](size_t i) const
    {
        if (d_->kind != Sequence || i >= d_->seq.size()) return Node();
        return d_->seq[i];
    }
    const Node operator[](int i) const { return (*this)[(size_t)i]; }

    template <class T> T as() const;
    const std::string& Scalar_() const { return d_->scalar; }

    // construction (used by the loader)
    static Node makeScalar(const std::string& s) { Node n; n.d_->kind = Scalar; n.d_->scalar = s; return n; }
    static Node makeSequence() { Node n; n.d_->kind = Sequence; return n; }
    static Node makeMap() { Node n; n.d_->kind = Map; return n; }
    void push_back(const Node& n) { d_->kind = Sequence; d_->seq.push_back(n); }
    void set(const std::string& k, const Node& n) { d_->kind = Map; d_->map[k] = n; }

private:
    struct Data {
        Kind kind = Undefined;
        std::string scalar;
        std::vector<Node> seq;
        std::map<std::string, Node> map;
    };
    std::shared_ptr<Data> d_;
};

namespace detail {
inline std::string trim(const std::string& s)
{
    size_t a = s.find_first_not_of(" \t\r\n");
    if (a == std::string::npos) return "";
    size_t b = s.find_last_not_of(" \t\r\n");
    return s.substr(a, b - a + 1);
}
inline std::string strip_comment(const std::string& s)
{
    bool in_single = false, in_double = false;
    for (size_t i = 0; i < s.size(); ++i) {
        char c = s[i];
        if (c == '\'' && !in_double) in_single = !in_single;
        else if (c == '"' && !in_single) in_double = !in_double;
        else if (c == '#' && !in_single && !in_double && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) return s.substr(0, i);
    }
    return s;
}
inline std::string unquote(const std::string& s)
{
    if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) return s.substr(1, s.size() - 2);
    return s;
}
inline Node parse_value(const std::string& text)
{
    std::string v = trim(text);
    if (!v.empty() && v.front() == '[') {
        if (v.back() != ']') throw ParserException("unterminated flow sequence: " + v);
        Node seq = Node::makeSequence();
        std::string inner = v.substr(1, v.size() - 2), item;
        std::stringstream ss(inner);
        while (std::getline(ss, item, ',')) {
            item = trim(item);
            if (!item.empty()) seq.push_back(Node::makeScalar(unquote(item)));
        }
        return seq;
    }
    return Node::makeScalar(unquote(v));
}
struct Line { int indent; std::string text; };

inline Node parse_block(const std::vector<Line>& lines, size_t& pos, int indent)
{
    if (pos >= lines.size()) return Node();
    if (lines[pos].text.compare(0, 2, "- ") == 0 || lines[pos].text == "-") {
        Node seq = Node::makeSequence();
        while (pos < lines.size() && lines[pos].indent == indent && (lines[pos].text.compare(0, 2, "- ") == 0 || lines[pos].text == "-")) {
            std::string rest = trim(lines[pos].text.substr(1));
            ++pos;
            if (rest.empty()) {
                if (pos < lines.size() && lines[pos].indent > indent) seq.push_back(parse_block(lines, pos, lines[pos].indent));
                else seq.push_back(Node::makeScalar(""));
            } else {
                seq.push_back(parse_value(rest));
            }
        }
        return seq;
    }
    Node map = Node::makeMap();
    while (pos < lines.size() && lines[pos].indent == indent) {
        const std::string& t = lines[pos].text;
        size_t colon = t.find(':');
        if (colon == std::string::npos) throw ParserException("expected 'key: value' in line: " + t);
        std::string key = unquote(trim(t.substr(0, colon)));
        std::string rest = trim(t.substr(colon + 1));
        ++pos;
        if (rest.empty()) {
            if (pos < lines.size() && lines[pos].indent > indent) map.set(key, parse_block(lines, pos, lines[pos].indent));
            else map.set(key, Node::makeScalar(""));
        } else {
            map.set(key, parse_value(rest));
        }
    }
    if (pos < lines.size() && lines[pos].indent > indent) throw ParserException("bad indentation near: " + lines[pos].text);
    return map;
}
}  // namespace detail

inline Node Load(const std::string& text)
{
    std::vector<detail::Line> lines;
    std::stringstream ss(text);
    std::string raw;
    while (std::getline(ss, raw)) {
        std::string s = detail::strip_comment(raw);
        if (detail::trim(s).empty() || detail::trim(s) == "---") continue;
        int indent = 0;
        while (indent < (int)s.size() && s[indent] == ' ') ++indent;
        if (indent < (int)s.size() && s[indent] == '\t') throw ParserException("tabs are not allowed for indentation");
        lines.push_back({indent, detail::trim(s)});
    }
    if (lines.empty()) return Node();
    size_t pos = 0;
    Node root = detail::parse_block(lines, pos, lines[0].indent);
    if (pos != lines.size()) throw ParserException("bad indentation near: " + lines[pos].text);
    return root;
}

inline Node LoadFile(const std::string& path)
{
    std::ifstream f(path.c_str());
    if (!f) throw BadFile(path);
    std::stringstream ss;
    ss << f.rdbuf();
    return Load(ss.str());
}

template <> inline std::string Node::as<std::string>() const
{
    if (d_->kind != Scalar) throw BadConversion("not a scalar");
    return d_->scalar;
}
template <> inline double Node::as<double>() const
{
    if (d_->kind != Scalar) throw BadConversion("not a scalar");
    char* end = nullptr;
    const double v = std::strtod(d_->scalar.c_str(), &end);
    if (end == d_->scalar.c_str() || *end != '\0') throw BadConversion("'" + d_->scalar + "' is not a number");
    return v;
}
template <> inline int Node::as<int>() const
{
    if (d_->kind != Scalar) throw BadConversion("not a scalar");
    char* end = nullptr;
    const long v = std::strtol(d_->scalar.c_str(), &end, 10);
    if (end == d_->scalar.c_str() || *end != '\0') throw BadConversion("'" + d_->scalar + "' is not an integer");
    return (int)v;
}
template <> inline bool Node::as<bool>() const
{
    if (d_->kind != Scalar) throw BadConversion("not a scalar");
    const std::string& s = d_->scalar;
    if (s == "true" || s == "True" || s == "TRUE" || s == "yes" || s == "on") return true;
    if (s == "false" || s == "False" || s == "FALSE" || s == "no" || s == "off") return false;
    throw BadConversion("'" + s + "' is not a boolean");
}

}  // namespace YAML
